"""CPU restatement of the action selection in main.py.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/code/main.py:158-233 (exploit path only; the epsilon-greedy
override at :211-217,:234-241 is random and stays in the caller).  `main.py` is
not runnable as published (undefined names), so this
restatement is anchored on the source text, not on an executed reference
("parity unpinned" for this one function; it is 30 lines of numpy argmax).
"""
import numpy as np


def select_action(gra_conf, suc_conf, gs_conf=None, is_ets=False, method="reinforcement"):
    """gra_conf/suc_conf: [N,R] float64; gs_conf: [N,N] upper triangle (others -100) or None.

    Returns dict(primitive, bestg_id, bests_id, bestgs_num, bestgs_g_id, bestgs_s_id,
    bestg_conf, bests_conf, bestgs_conf).
    """
    gra_conf = np.asarray(gra_conf, dtype=np.float64)
    suc_conf = np.asarray(suc_conf, dtype=np.float64)
    n = gra_conf.shape[0]
    gnu_best = gra_conf.max(axis=1)                       # main.py:167
    gro_best = gra_conf.argmax(axis=1).astype(int)        # main.py:168
    sro_best = suc_conf.argmax(axis=1).astype(int)
    bestg_conf, bests_conf = gra_conf.max(), suc_conf.max()          # main.py:170
    bestg_id = np.unravel_index(np.argmax(gra_conf), gra_conf.shape)  # main.py:172 (first max wins)
    bests_id = np.unravel_index(np.argmax(suc_conf), suc_conf.shape)
    bestgs_conf = 0
    bestgs_num, bestgs_g_id, bestgs_s_id = [], [], []
    if is_ets and n > 1:
        gs = np.asarray(gs_conf, dtype=np.float64)
        bestgs_conf = gs.max()                                        # main.py:194
        bestgs_num = np.unravel_index(np.argmax(gs), gs.shape)        # main.py:195
        a, b = bestgs_num
        if gnu_best[a] > gnu_best[b]:                                 # main.py:196-201
            bestgs_g_id, bestgs_s_id = [a, gro_best[a]], [b, sro_best[b]]
        else:
            bestgs_g_id, bestgs_s_id = [b, gro_best[b]], [a, sro_best[a]]
    primitive = "grasp"                                               # main.py:205
    if not is_ets or n == 1:
        if bests_conf > bestg_conf:                                   # main.py:208
            primitive = "suction"
    else:
        g2 = 2 * bestgs_conf if method == "reactive" else bestgs_conf  # main.py:221-233
        if bests_conf > max(bestg_conf, g2):
            primitive = "suction"
        elif g2 > max(bests_conf, bestg_conf):
            primitive = "grasp_then_suction"
    return {"primitive": primitive, "bestg_id": tuple(int(v) for v in bestg_id),
            "bests_id": tuple(int(v) for v in bests_id),
            "bestgs_num": tuple(int(v) for v in bestgs_num),
            "bestgs_g_id": [int(v) for v in bestgs_g_id], "bestgs_s_id": [int(v) for v in bestgs_s_id],
            "bestg_conf": float(bestg_conf), "bests_conf": float(bests_conf),
            "bestgs_conf": float(bestgs_conf)}
