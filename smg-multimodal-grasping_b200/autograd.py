"""Grad-enabled branch of the nets' forward (code/models.py:513-586): autograd bridge to smg_qbackward."""
import torch


def q_forward_with_grad(model, input_depth_data, m_input_depth_data, style, specific_rotation):
    raise NotImplementedError(
        "smg_b200: the grad-enabled forward (trainer.backprop) needs smg_qforward_train/smg_qbackward, "
        "which this build does not provide yet; there is no PyTorch fallback by design.")
