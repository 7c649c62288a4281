"""world_size-2 gloo tests of the batch scatter / waveform gather host logic.
The engine itself is CUDA-only, so a deterministic CPU stub with the infer() signature stands in."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wetts_b200.dist import plan_shards


class StubNet:
    """o[b,0,n] = sum(ids[b,:len]) + sid[b] + n/1000 (+ noise_z[b,0,n // U] when given) for n < 2*len*U;
    y_lengths = 2*len."""
    U = 4

    def infer(self, x, x_lengths, sid=None, noise_z=None, **kw):
        B, Tx = x.shape
        ylen = 2 * x_lengths
        Ty = int(ylen.max())
        y_mask = (torch.arange(Ty)[None, :] < ylen[:, None]).float()[:, None]
        m = (torch.arange(Tx)[None, :] < x_lengths[:, None])
        base = (x * m).sum(dim=1).float() + (0 if sid is None else sid.float())
        o = base[:, None, None] + torch.arange(Ty * self.U).float()[None, None, :] / 1000.0
        if noise_z is not None:   # per-utterance injected noise must follow its utterance to whichever rank gets it
            o = o + noise_z[:, :1, :Ty].repeat_interleave(self.U, dim=2)
        return o, None, y_mask, None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wetts_b200.dist import sharded_infer
    gen = torch.Generator().manual_seed(3)
    B, Tx = 7, 9
    lens = torch.tensor([9, 3, 7, 5, 9, 2, 6])
    x = torch.randint(1, 50, (B, Tx), generator=gen) * (torch.arange(Tx)[None, :] < lens[:, None])
    sid = torch.randint(0, 3, (B,), generator=gen)
    net = StubNet()
    noise_z = torch.randn(B, 3, 2 * Tx, generator=gen)
    out = sharded_infer(net, x if rank == 0 else None, lens if rank == 0 else None, sid if rank == 0 else None,
                        torch.device("cpu"), hop_upsample=StubNet.U, noise_z=noise_z if rank == 0 else None)
    if rank == 0:
        ref, _, ym, _ = net.infer(x, lens, sid, noise_z=noise_z)
        ok = len(out) == B
        for i in range(B):
            n = int(ym[i].sum()) * StubNet.U
            ok = ok and out[i].shape[0] == n and torch.equal(out[i], ref[i, 0, :n])
        ret["ok"] = bool(ok)
    dist.destroy_process_group()


def test_plan_shards_balances_and_covers():
    lens = [128, 64, 100, 90, 70, 128, 10, 55, 31]
    plan = plan_shards(lens, 4)
    flat = sorted(i for s in plan for i in s if i >= 0)
    assert flat == list(range(len(lens)))
    assert len({len(s) for s in plan}) == 1
    sums = [sum(lens[i] for i in s if i >= 0) for s in plan]
    assert max(sums) - min(sums) <= max(lens)


def test_sharded_equals_unsharded_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
