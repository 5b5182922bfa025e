"""CPU tests of the host-side bookkeeping that decides when packed weights are stale (engine.sync_weights)."""
import copy

import torch


def test_param_signature_tracks_versions_epochs_and_moves():
    from smg_b200 import engine as E
    from smg_b200 import models
    torch.manual_seed(0)
    m = models.reinforcement_net(False)
    params = E.trunk_param_list(m.grasp_depth_trunk)
    assert len(params) == 362
    s0 = E._param_signature(m, 0, params)
    assert s0 == E._param_signature(m, 0, params)
    with torch.no_grad():
        params[100].add_(1.0)                                  # any in-place torch op bumps the version counter
    s1 = E._param_signature(m, 0, params)
    assert s1 != s0
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.load_state_dict(sd)                                      # copies in place: versions move, addresses stay
    s2 = E._param_signature(m, 0, params)
    assert s2 != s1 and s2[3:6] == s1[3:6]
    assert E._param_signature(m, 1, params) != s2              # the library's own in-place update counter
    moved = [torch.nn.Parameter(p.detach().clone()) for p in params]
    assert E._param_signature(m, 0, moved)[3:6] != s2[3:6]     # a moved model changes the probed addresses


def test_weight_epochs_are_per_trunk_and_head_and_copied_with_the_model():
    from smg_b200 import engine as E
    from smg_b200 import models
    m = models.reinforcement_net(False)
    assert getattr(m, "_smg_epoch", None) in (None, {})
    E.bump_weight_epoch(m, 0)                                  # style 0 -> grasp trunk + grasp head
    tid, hid = E.STYLE_ROUTE[0]
    assert m._smg_epoch == {("t", tid): 1, ("h", hid): 1}
    E.bump_weight_epoch(m, 0)
    E.bump_weight_epoch(m, 1)
    t1, h1 = E.STYLE_ROUTE[1]
    assert m._smg_epoch[("t", tid)] == 2 and m._smg_epoch[("t", t1)] == 1 and m._smg_epoch[("h", h1)] == 1
    target = copy.deepcopy(m)                                  # model_target: its own counters from then on
    E.bump_weight_epoch(m, 0)
    assert target._smg_epoch[("t", tid)] == 2 and m._smg_epoch[("t", tid)] == 3
