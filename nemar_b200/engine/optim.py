"""Flat-buffer Adam + gradient bucket (replaces torch.optim.Adam, reference models/nemar_model.py:124-141).

All parameters of one optimizer live in ONE contiguous fp32 buffer (params / grads / exp_avg / exp_avg_sq),
so the update is a single kernel launch and the data-parallel exchange is a single NCCL all-reduce on the
gradient bucket.  nn.Parameter objects stay what the reference exposes (same names,
same shapes): their .data/.grad are re-pointed at slices of the flat buffers.
"""
import torch

from . import functional as F


class FlatAdam:
    def __init__(self, params, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
        self.params = [p for p in params]
        assert len(self.params) > 0
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4   # keep every slice 16-byte aligned
        self.numel = total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                n = p.numel()
                self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + n].view(p.shape)
                p.grad = self.flat_g[o:o + n].view(p.shape)
        self.offsets = offs
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # device-resident step counter (CUDA-graph replay)
        # torch.optim-like surface used by BaseModel (schedulers / lr printing)
        self.param_groups = [{"params": self.params, "lr": lr, "betas": betas, "eps": eps}]
        self.grad_hook = None   # set by the data-parallel wrapper: called on flat_g before the update
        F.bump_weights_epoch()

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        for p, o in zip(self.params, self.offsets):   # re-attach views autograd may have replaced
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)

    def step(self):
        F.join_wgrad_stream()          # weight gradients run on their own stream: the bucket is complete after this
        grad_scale = 1.0
        if self.grad_hook is not None:
            grad_scale = self.grad_hook(self.flat_g)
        self.step_count += 1
        lr = self.param_groups[0]["lr"]
        # the step number lives in device memory so that a captured step replays with the right bias correction
        F.adam_step_dev(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, lr, self.betas[0], self.betas[1],
                        self.eps, self.step_dev, grad_scale)
        F.refresh_packs(self.flat_p)      # ONE launch re-packs the bf16 operand layouts of every layer just updated

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.detach().cpu().clone(),
                "exp_avg_sq": self.exp_avg_sq.detach().cpu().clone(), "lr": self.param_groups[0]["lr"],
                "layout": [int(p.numel()) for p in self.params]}

    def load_state_dict(self, sd):
        if "layout" in sd and list(sd["layout"]) != [int(p.numel()) for p in self.params]:
            raise ValueError("optimizer state was saved for another parameter layout")
        self.step_count = int(sd["step"])
        self.step_dev.fill_(self.step_count)
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0]["lr"] = float(sd.get("lr", self.lr))
