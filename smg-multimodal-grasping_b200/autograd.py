"""Grad-enabled branch of the nets' forward (code/models.py:513-586) as a torch.autograd.Function
over smg_qforward_train / smg_qbackward.

The reference builds an autograd graph through two DenseNet passes and the head; `trainer.backprop`
(code/trainer.py:338-351) then indexes `model.gra_prob[0,0,0,0]`, forms the loss and calls `.backward()`,
which must populate `.grad` on the module's own nn.Parameters (read by torch.optim.Adam,
code/trainer.py:99,383).  Here the whole Q pass is ONE autograd node: its inputs are the 362 trunk
parameters + 6 head parameters the sample touches, its backward asks the CUDA library for all 368
gradients at once.  Parameters of the other two trunks / heads get no gradient, exactly as in the
reference (Adam skips `grad is None`).
"""
import torch

from . import engine as _engine


class _QPass(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, style, scene, mask, rot, nrot, n_trunk, *params):
        q, mean, var = eng.qforward_train(style, scene, mask, rot, nrot)
        ctx.eng = eng
        ctx.pass_id = eng.train_pass_id()
        ctx.n_trunk = n_trunk
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.mark_non_differentiable(mean, var)
        return q, mean, var

    @staticmethod
    def backward(ctx, dq, _dm, _dv):
        eng = ctx.eng
        if eng.train_pass_id() != ctx.pass_id:
            raise RuntimeError("smg_b200: the activations saved by this grad-enabled forward were overwritten by another "
                               "forward on the same model (or its engine was re-created) before backward(); only one "
                               "grad-enabled pass may be in flight per model")
        grads = [torch.empty(s, dtype=torch.float32, device=eng.device) for s in ctx.shapes]
        eng.qbackward(dq.contiguous().view(-1), grads[:ctx.n_trunk], grads[ctx.n_trunk:])
        return (None,) * 7 + tuple(grads)


def q_forward_with_grad(model, input_depth_data, m_input_depth_data, style, specific_rotation):
    """(is_volatile=False, specific_rotation=r): returns a [1,C,1,1] tensor with grad_fn and stores it in
    model.gra_prob / suc_prob / gs_prob like the reference (code/models.py:539,561,584)."""
    rot = 0 if style == 2 else int(specific_rotation)      # ES is pinned to rotation 0 (code/models.py:567)
    eng = model._engine(2, style)
    tid, hid = _engine.STYLE_ROUTE[int(style)]
    trunk = getattr(model, _engine.TRUNK_ATTRS[tid])
    head = getattr(model, _engine.HEAD_ATTRS[hid])
    tparams = _engine.trunk_param_list(trunk)
    hparams = _engine.head_param_list(head)
    for p in tparams + hparams:
        if not p.is_cuda:
            raise RuntimeError("smg_b200: training needs the model on the GPU (call model.cuda())")
    scene = input_depth_data.reshape(3, 640, 640)
    mask = m_input_depth_data.reshape(3, 640, 640)
    q, mean, var = _QPass.apply(eng, style, scene, mask, rot, model.gnum_rotations, len(tparams), *tparams, *hparams)
    if model.update_running_stats:
        model._apply_running_stats(trunk, mean, var, [0, 1])   # trunk(scene) then trunk(mask)
        model._apply_head_running_stats(head, trunk, var, [(0, 1)], eng.head_bn_stats(1))
    out = q.view(1, model.N_OUT, 1, 1)
    model.gra_prob, model.suc_prob, model.gs_prob = [], [], []
    setattr(model, {0: "gra_prob", 1: "suc_prob", 2: "gs_prob"}[int(style)], out)
    return out
