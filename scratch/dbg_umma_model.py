import sys
sys.path[:0] = ["/root/repo", "/root/repo/graphcast-lite_b200", "/root/repo/tests", "/root/repo/oracle/pyg_shim", "/root/repo/oracle/trimesh_shim"]
import torch, copy
from test_model_gpu import _models
from gcl_b200.train import Trainer
from gcl_b200 import _cabi
from oracle import model as om
from helpers import rel_err
lib = _cabi.load()
name = sys.argv[1]
mine, ref, cfg, nlat, nlon = _models(name)
G = nlat*nlon; F, T = cfg["data"]["num_features_used"], cfg["data"]["obs_window_used"]
gen = torch.Generator().manual_seed(3)
X, y = torch.randn(1, G, T*F, generator=gen), torch.randn(1, G, F, generator=gen)
ref64 = copy.deepcopy(ref).double()
for attr in ("init_grid_features","init_mesh_features"): setattr(ref64, attr, getattr(ref64, attr).double())
lc = om.training_loss(ref, X, y, 1, om.lat_weights(nlat, nlon)); lc.backward()
l64 = om.training_loss(ref64, X.double(), y.double(), 1, om.lat_weights(nlat, nlon).double()); l64.backward()
tr = Trainer(mine, nlat, nlon, ar_steps=1)
grads = {}
for mode in (1, 0):
    lib.gcl_set_dense_mode(mode)
    tr.zero_grad(); l = tr.loss(X.cuda(), y.cuda()); l.backward()
    grads[mode] = {k: p.grad.detach().cpu().clone() for k, p in mine.named_parameters()}
p32 = dict(ref.named_parameters()); p64 = dict(ref64.named_parameters())
rows = []
for k in grads[0]:
    if p64[k].grad is None: continue
    t = p64[k].grad.float()
    rows.append((rel_err(grads[0][k], t), rel_err(grads[1][k], t), rel_err(p32[k].grad, t), k))
rows.sort(reverse=True)
print("worst by umma-vs-f64:   umma      ffma     cpu32")
for r in rows[:8]: print("  %-48s %.2e %.2e %.2e" % (r[3], r[0], r[1], r[2]))
