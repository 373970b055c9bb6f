"""Drop-in mirror of the reference's ``model`` package for the guided-sampling path.

Same import surface (``from model.diffusion import DDPMTrainer, DDPMSampler, DDIMSampler``,
``from model.decoder import MipConditionalDecorder``, ``from model.cmsp import CMSP``), same
constructor signatures and ``state_dict`` layouts as /root/reference/model/*.py, but every forward
computation on the sampling path runs in the sm_100a kernels of libdip_b200.so.
"""
