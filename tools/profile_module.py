"""Where does a whole GroupBimodalCSRPool training step spend its time? (developer tool)
python tools/profile_module.py [N] [views] [C]"""
import sys
import torch
sys.path.insert(0, ".")
from deepviewagg_b200.modules.multimodal.pooling import GroupBimodalCSRPool

_args = [a for a in sys.argv[1:] if not a.startswith("--")]
N, v, C = (int(a) for a in (_args[:3] + [160000, 8, 64][len(_args):]))
USE_MOD = "--no-mod" not in sys.argv          # the Group-pool YAMLs of the reference set use_mod: False
dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(0)
counts = torch.poisson(torch.full((N,), float(v), device=dev), generator=gen).long()
ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), counts.cumsum(0)])
V = int(ptr[-1])
m = GroupBimodalCSRPool(in_map=8, in_mod=C, num_groups=4, use_num=True, use_mod=USE_MOD).to(dev).train()
x_mod = torch.randn(V, C, device=dev, requires_grad=True)
x_map = torch.rand(V, 8, device=dev)
w = torch.randn(N, C, device=dev)


def step():
    out = m(None, x_mod, x_map, ptr)
    (out * w).sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    step()
b.record()
torch.cuda.synchronize()
print(f"N={N} v={v} C={C} V={V}: {a.elapsed_time(b) / 10:.3f} ms/step  -> {N / (a.elapsed_time(b) / 10) / 1e3:.2f} Mpoints/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
