"""Block-by-block diagnostic on the GPU box: prints the error of every stage against the
reference-generated fixtures without stopping at the first mismatch."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wetts_b200
from tests.golden_util import CASES, load_case, rel_rms_err


def rep(tag, a, b):
    a = a.detach().cpu()
    try:
        print(f"  {tag:28s} rel_rms_err={rel_rms_err(a, b):.3e}  max|ref|={float(b.abs().max()):.3e} nan={bool(torch.isnan(a).any())}", flush=True)
    except Exception as e:
        print(f"  {tag:28s} FAILED {e}", flush=True)


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    from wetts_b200 import _lib
    tc = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    _lib.check(_lib.load().wetts_set_option(b"tensor_cores", tc))
    fused = int(os.environ.get("WETTS_FUSED_RB", "1"))
    _lib.check(_lib.load().wetts_set_option(b"fused_resblock", fused))
    print("tensor_cores =", tc, "fused_resblock =", fused, flush=True)
    cases = sys.argv[2:] if len(sys.argv) > 2 else CASES
    for name in cases:
        print("case", name, flush=True)
        try:
            hps, sd, g, t = load_case(name)
            t0 = time.time()
            net = wetts_b200.build_model(hps, int(g["n_vocab"]), int(g["n_speakers"]), sd, "cuda")
            torch.cuda.synchronize()
            print(f"  build+finalize {time.time()-t0:.2f}s", flush=True)
            dev = net.device
            ns, ls, nsw = [float(v) for v in g["scales"]]
            gvec = net.emb_g(t["sid"])[:, :, None] if int(g["n_speakers"]) > 0 else None
            h, m, logs, x_mask = net.enc_p(t["x"], t["x_lengths"])
            rep("enc_p.h", h, t["h"]); rep("enc_p.m", m, t["m_p_tx"]); rep("enc_p.logs", logs, t["logs_p_tx"])
            if net.use_sdp:
                logw = net.dp(t["h"].to(dev), x_mask, g=gvec, reverse=True, noise_scale=nsw, noise=t["noise_w"])
            else:
                logw = net.dp(t["h"].to(dev), x_mask, g=gvec)
            rep("dp.logw(given ref h)", logw, t["logw"])
            Ty = t["z"].shape[2]
            y_mask = (torch.arange(Ty)[None, :] < t["y_lengths"][:, None]).float()[:, None].to(dev)
            z = net.flow(t["z_p"].to(dev), y_mask, g=gvec, reverse=True)
            rep("flow(given ref z_p)", z, t["z"])
            o = net.dec(t["z"].to(dev) * y_mask, g=gvec)
            rep("dec(given ref z)", o, t["o"])
            o, attn, ym, (z, z_p, m_p, logs_p) = net.infer(t["x"], t["x_lengths"], t["sid"], ns, ls, nsw,
                                                           noise_w=t["noise_w"], noise_z=t["noise_z"], durations=t["w_ceil"])
            print("  y_lengths", net.last_y_lengths.tolist(), "ref", t["y_lengths"].tolist(), flush=True)
            rep("e2e z_p", z_p, t["z_p"]); rep("e2e z", z, t["z"]); rep("e2e o", o, t["o"])
            o2, *_ = net.infer(t["x"], t["x_lengths"], t["sid"], ns, ls, nsw, noise_w=t["noise_w"], noise_z=t["noise_z"])
            print("  own durations y_lengths", net.last_y_lengths.tolist(), flush=True)
            zt = t["z"].transpose(1, 2).contiguous()
            o3 = net.export_decoder_forward(zt, t["sid"])
            o_ref_nomask = None
            print("  decoder contract out", tuple(o3.shape), "finite", bool(torch.isfinite(o3).all()), flush=True)
        except Exception:
            traceback.print_exc()
            sys.stdout.flush()


if __name__ == "__main__":
    main()
