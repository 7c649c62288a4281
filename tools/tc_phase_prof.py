"""Per-phase cycle decomposition of conv1d_tc_kernel on selected generator convs (debug hook)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, wetts_b200
from wetts_b200 import synth, _lib
from wetts_b200.hparams import builtin_config
lib = _lib.load()
lib.wetts_debug_set_tc_profile.argtypes = [C.c_void_p]
hps = builtin_config("multilingual_v3")
sd = synth.make_state_dict(hps.model, 256, 2, seed=1234)
net = wetts_b200.build_model(hps, 256, 2, sd, "cuda")
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 754
z = torch.randn(B, 192, T, device="cuda"); g = net.emb_g(torch.zeros(B, dtype=torch.long))[:, :, None]
for _ in range(2): net.dec(z, g=g)
torch.cuda.synchronize()
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.wetts_debug_set_tc_profile(C.c_void_p(buf.data_ptr()))
net.dec(z, g=g); torch.cuda.synchronize()
lib.wetts_debug_set_tc_profile(None)
v = buf.cpu().tolist()
names = ["stage", "a_free wait", "acc wait", "epilogue", "item total", "-", "-", "-"]
print("stager warp0 (sum over CTAs & all TC launches of the generator), Mcycles:")
for n, x in zip(names[:5], v[:5]): print(f"  {n:12s} {x/1e6:10.1f}")
print("MMA warp, Mcycles:")
tiles = v[15] >> 40
for n, x in zip(["acc wait", "epilogue", "item total", "a_full wait", "issue", "b_full wait"], [v[10], v[11], v[12], v[13], v[14], v[15] & ((1 << 40) - 1)]):
    print(f"  {n:12s} {x/1e6:10.1f}")
print("tiles issued", tiles, " issue cycles/tile", v[14] / max(tiles, 1), " a_full wait cycles/tile", v[13] / max(tiles, 1))
