"""sdumc_b200 - B200-native (sm_100a) hot path of WarmCongee/SDUMC: the UMC model, the full / text-missing
self-distillation train step and the two-pass scoring path behind the reference's Python API.

    sdumc_b200.build      nvcc build of csrc/ -> libsdumc_b200.so (C ABI: include/sdumc_b200.h)
    sdumc_b200._lib       ctypes binding generated from the header
    sdumc_b200.ops        torch front ends of the C-ABI operators
    sdumc_b200.engine     forward / backward orchestration of the kernels
    sdumc_b200.model      drop-in nn.Module (toolkit.models.get_models) and autograd bridge
    sdumc_b200.losses     MSELoss / RMSELoss / RnCLoss (toolkit.utils.loss)
    sdumc_b200.trainer    fused train step (both passes, loss, backward, Adam; CUDA-graph replay), scoring
    sdumc_b200.dp         data-parallel exchanges (NCCL / gloo)
    sdumc_b200.dataset    feature stores (pinned host, HBM-resident) and the reference's collate rule
    sdumc_b200.cli        main_frame_val_text_missing(.py|_inference.py)

Importing the package does not load the CUDA library; the operators raise SdumcError when it is missing or
when no CUDA device is present (there is no CPU fallback)."""

__version__ = "0.1.0"
