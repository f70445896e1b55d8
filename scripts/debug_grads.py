"""GPU debug: compare intermediate gradients of the T/R phase (no D update) between engine (fp32) and oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nemar_oracle as O
from tests import helpers as H
from nemar_b200.engine import functional as F

name = sys.argv[1] if len(sys.argv) > 1 else "c1_affine64"
model, cfg, (T, R, Ds), (A, B) = H.build_case(name)
model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
model.forward()
eng = {}
for k in ("fake_B", "registered_real_A", "fake_TR_B", "fake_RT_B"):
    t = getattr(model, k)
    t.retain_grad()
    eng[k] = t
model.set_requires_grad([model.netD, *model.netD_multiresolution], False)
model.optimizer_TR.zero_grad()
model.backward_T_and_R()
torch.cuda.synchronize()

st = O.OracleStep(cfg, T, R, Ds)
o = st.forward(A, B)
for k in ("fake_B", "registered_real_A", "fake_TR_B", "fake_RT_B", "grid"):
    o[k].retain_grad()
Dn = [dict((k, v.detach()) for k, v in d.items()) for d in st.Ds]
loss = cfg.lambda_recon * torch.nn.functional.l1_loss(o["fake_TR_B"], B) + cfg.lambda_recon * torch.nn.functional.l1_loss(o["fake_RT_B"], B) \
    + cfg.lambda_gan * st._gan(A, o["fake_TR_B"], True, Dn) + cfg.lambda_gan * st._gan(A, o["fake_RT_B"], True, Dn) + cfg.lambda_smooth * o["reg"]
loss.backward()

def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20)), float(b.norm())

for k in ("fake_TR_B", "fake_RT_B", "registered_real_A", "fake_B"):
    print("value %-20s rel %.3e" % (k, rel(eng[k], o[k])[0]), "  grad rel %.3e |ref| %.3e" % rel(eng[k].grad, o[k].grad))
for tag, net, sd in (("T", model.netT, st.T), ("R", model.netR, st.R)):
    worst = sorted(((rel(p.grad, sd[k].grad)[0], k) for k, p in net.named_parameters() if k.endswith("weight")), reverse=True)[:4]
    print(tag, "worst weight grads:", ["%s %.2e" % (k, r) for r, k in worst])
