"""world_size-2 NCCL test of the serving-side sharding (wetts_b200/dist.py) on two real GPUs: rank 0 holds the batch,
ids (and the injected per-utterance noise) are scattered over NCCL, both ranks synthesise their shard with the CUDA
engine, waveforms are gathered back to rank 0 and must be BIT-identical to the unsharded run on rank 0's GPU (SURVEY.md 8e parity caveat:
explicit per-utterance noise), except inside the vocoder's receptive field of the end of a shard's longest utterance
(the reference model's own dependence on batch padding; see the comment in the worker).  Skipped on boxes with fewer than two GPUs
(run it with `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    import wetts_b200
    from wetts_b200 import synth
    from wetts_b200.dist import sharded_infer
    from wetts_b200.hparams import builtin_config
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    hps = builtin_config("multilingual_v3")
    n_vocab, n_spk = 64, 2
    sd = synth.make_state_dict(hps.model, n_vocab, n_spk, seed=1234)
    net = wetts_b200.build_model(hps, n_vocab, n_spk, sd, dev)
    gen = torch.Generator().manual_seed(3)
    B, Tx = 7, 24
    lens = torch.tensor([24, 9, 17, 13, 24, 5, 20])
    x = torch.randint(1, n_vocab, (B, Tx), generator=gen) * (torch.arange(Tx)[None, :] < lens[:, None])
    sid = torch.randint(0, n_spk, (B,), generator=gen)
    # 3..6 frames per phoneme: every shard holds a 24-phoneme utterance, so every shard and the unsharded batch have
    # Ty >= 72 frames and take the same kernel route (flow convs switch from the fp32 SIMT kernel to the tensor pipe at
    # T >= 64; bit-equality between routes is not a property of the engine)
    dur = torch.randint(3, 7, (B, 1, Tx), generator=gen).float() * (torch.arange(Tx)[None, None, :] < lens[:, None, None])
    Tmax = int(dur.sum(-1).max())
    noise_z = torch.randn(B, 192, Tmax, generator=gen)
    kw = dict(noise_scale=0.667, length_scale=1.0, noise_scale_w=0.8, return_attn=False)
    on0 = rank == 0
    out = sharded_infer(net, x.to(dev) if on0 else None, lens.to(dev) if on0 else None, sid.to(dev) if on0 else None, dev,
                        noise_z=noise_z if on0 else None, durations=dur if on0 else None, **kw)
    if on0:
        o, _, ym, _ = net.infer(x, lens, sid, noise_z=noise_z, durations=dur, **kw)
        # Bit-identical, with the one exception the reference model has too: HiFi-GAN does not mask, so the last samples of
        # an utterance depend on whether it is followed by batch padding ("processed zeros") or by the end of the tensor.
        # The longest utterance of a shard loses its padding when it was not the longest of the whole batch, and may
        # differ inside the vocoder's receptive field of its end (< 16 frames); everything before that is bit-identical
        # (measured: tools/diag_batch_invariance.py, z and z_p are identical in every case).
        ok = len(out) == B
        why = []
        margin, tail_diffs = 16 * 256, 0
        for i in range(B):
            n = int(ym[i].sum()) * 256
            if out[i].shape[0] != n:
                why.append(f"utterance {i}: {out[i].shape[0]} samples, expected {n}")
            elif not torch.equal(out[i], o[i, 0, :n]):
                if torch.equal(out[i][:n - margin], o[i, 0, :n - margin]):
                    tail_diffs += 1
                else:
                    why.append(f"utterance {i}: differs before the last {margin} samples, max |diff| "
                               f"{float((out[i] - o[i, 0, :n]).abs().max()):.3e}")
        if tail_diffs > world:
            why.append(f"{tail_diffs} utterances differ at the tail; at most one per shard can")
        ret["ok"] = bool(ok and not why)
        ret["why"] = "; ".join(why)
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_sharded_equals_unsharded_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get("ok") is True, ret.get("why")
