"""CPU restatement of `Trainer.get_label_value`.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/code/trainer.py:212-274.  Reactive: label 0 = success, 1 = failure (ES succeeds only with reward
2.5).  Reinforcement: y = r + gamma * Q_target(s', a*) where a* is the CURRENT step's best action of the exploited
primitive evaluated by the target net at its best rotation; the future term is 0 when everything failed or the table is
cleared.  Pinned on tests/golden/golden_r02.json (15 cases recorded from the unmodified reference).
"""
import numpy as np
import torch

from . import qnet


def label_value(method, target_sd, args, depth_heightmap, mask_depth, gamma=0.5, num_rotations=1, mean=0.01, std=0.03):
    prim = args["primitive_action"]
    reward = {"suction": args["suction_success"], "grasp": args["grasp_success"],
              "grasp_then_suction": args["gs_success"]}[prim]
    if method == "reactive":                                                     # trainer.py:216-234
        ok = (reward == 2.5) if prim == "grasp_then_suction" else bool(reward)
        return (0 if ok else 1), reward
    n = args["objects_number"]
    s, g, gs = args["suction_success"], args["grasp_success"], args["gs_success"]
    if s == 0 and g == 0 and gs == 0:                                           # trainer.py:248-249
        future = 0.0
    elif (n == 1 and s == 1) or (n == 1 and g == 1) or (n == 2 and gs == 2.5):  # trainer.py:250-251
        future = 0.0
    else:                                                                        # trainer.py:259-270
        act = args["exploit_action"]
        if act == "grasp":
            style, ids, rot = 0, [args["bestg_id"][0]], args["bestg_id"][1]
        elif act == "suction":
            style, ids, rot = 1, [args["bests_id"][0]], args["bests_id"][1]
        else:
            style, ids, rot = 2, [args["bestgs_g_id"][0], args["bestgs_s_id"][0]], args["bestgs_g_id"][1]
        m = np.asarray(depth_heightmap, np.float64) * sum(np.asarray(mask_depth[i], np.float64) for i in ids)
        with torch.no_grad():
            q = qnet.model_forward(target_sd, qnet.preprocess(depth_heightmap, mean, std), qnet.preprocess(m, mean, std),
                                   style, True, rot, gnum_rotations=num_rotations, snum_rotations=num_rotations)
        future = float(q.view(-1)[0])
    return reward + gamma * future, reward                                       # trainer.py:271
