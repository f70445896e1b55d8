"""Engine-wide settings (set once from the command line, read by every layer)."""
import torch


class EngineConfig:
    def __init__(self):
        self.precision = "bf16"      # storage dtype of activations / packed weights: "bf16" | "fp32"
        self.conv_engine = "auto"    # "auto": tcgen05 where supported, else generic; "generic": CUDA cores only
        self.dropout_seed = 0x5EED
        # k7 head / tail of the generator: column taps as channels + a 7 x 1 tensor-core conv (True), or all 49 taps as
        # channels + a 1 x 1 conv (False: round 1's formulation, 5x the intermediate bytes)
        import os
        self.k7_xtaps = os.environ.get("NEMAR_K7_XTAPS", "1") != "0"
        # weight gradients on their own stream: a conv's wgrad is needed only by the optimizer, so it leaves the critical
        # path (dgrad chain) and fills the SMs the latency-bound passes leave idle
        self.wgrad_stream = os.environ.get("NEMAR_WGRAD_STREAM", "1") != "0"
        self.step_dev = None         # device-resident step counter (int64[1]) read by the dropout kernels
        self.dropout_calls = 0       # dropout call sites seen since the step began (-> a distinct salt per site)

    @property
    def dtype(self):
        return torch.bfloat16 if self.precision == "bf16" else torch.float32


CONFIG = EngineConfig()


def step_counter(device):
    """The device-resident step counter (created on first use)."""
    if CONFIG.step_dev is None or CONFIG.step_dev.device != torch.device(device):
        CONFIG.step_dev = torch.zeros(1, dtype=torch.int64, device=device)
    return CONFIG.step_dev


def configure(precision=None, conv_engine=None):
    if precision is not None:
        assert precision in ("bf16", "fp32")
        CONFIG.precision = precision
    if conv_engine is not None:
        assert conv_engine in ("auto", "generic")
        CONFIG.conv_engine = conv_engine
    return CONFIG
