"""smg-b200: B200-native (sm_100a) grasp-affordance hot path of SMG-multimodal-grasping.

Host side mirrors the reference's Python surface:
  models.reinforcement_net / models.reactive_net   code/models.py:301-586 / :15-296
  trainer.Trainer                                  code/trainer.py:17-384
  utils.get_heightmap                              code/utils.py:38-68
  NMS.py_cpu_nms                                   code/NMS.py:8-59
All device work is done by the C-ABI library `csrc/libsmg_b200.so`
(declared in include/smg_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
