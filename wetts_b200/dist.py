"""Batch-sharded multi-GPU inference: one process per GPU, weights replicated,
utterances dealt to ranks, no collective on the data path.  The only communication is the
"trivial batch scatter/gather" the north_star names: rank 0 holds the request batch, scatters
ids / lengths / speaker ids (KBs), every rank synthesises its shard, and the waveforms are
gathered back to rank 0 (NCCL over NVLink when the tensors are CUDA tensors; the same code runs
on gloo/CPU tensors for the host-logic tests).

The reference has no inference-time parallelism at all (inference.py:82-100 is batch 1 on one
device); its only sharding logic is the training-time DistributedBucketSampler
(wetts/vits/data_utils.py:228-346), whose length-balancing idea plan_shards() follows.
"""
import torch
import torch.distributed as dist


def plan_shards(x_lengths, world_size):
    """Deal utterances to ranks, longest first, so every rank gets a near-equal sum of phoneme
    counts (durations are unknown before the duration predictor has run; the phoneme count is
    the available proxy).  Returns a list (per rank) of index lists into the batch; every rank
    gets the same number of slots (the last ones may be padding = -1) so collectives stay regular."""
    lens = [int(v) for v in x_lengths]
    order = sorted(range(len(lens)), key=lambda i: (-lens[i], i))
    shards = [[] for _ in range(world_size)]
    # snake (boustrophedon) deal: 0..W-1, W-1..0, ...
    for k, idx in enumerate(order):
        rnd, pos = divmod(k, world_size)
        r = pos if rnd % 2 == 0 else world_size - 1 - pos
        shards[r].append(idx)
    per = max(len(s) for s in shards) if shards else 0
    return [s + [-1] * (per - len(s)) for s in shards]


def scatter_batch(x, x_lengths, sid, device, group=None, src=0):
    """Rank `src` passes the full batch (x int64[B,Tx], x_lengths int64[B], sid int64[B] or None);
    other ranks pass None.  Returns this rank's (x, x_lengths, sid, index_map): index_map[i] is
    the position of local utterance i in the original batch (-1 = padding slot, length 0 -> the
    engine still synthesises >= 1 frame for it; gather_audio drops it)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == src:
        plan = plan_shards(x_lengths.tolist(), world)
        meta[0], meta[1], meta[2] = len(plan[0]), x.shape[1], 0 if sid is None else 1
    dist.broadcast(meta, src, group=group)
    per, Tx, has_sid = int(meta[0]), int(meta[1]), bool(meta[2])
    # one packed int64 buffer per rank: [per, Tx + 3] = ids | length | sid | original index
    mine = torch.empty(per, Tx + 3, dtype=torch.int64, device=device)
    chunks = None
    if rank == src:
        xs, ls = x.to(device), x_lengths.to(device)
        ss = sid.to(device) if has_sid else torch.zeros_like(ls)
        chunks = []
        for shard in plan:
            idx = torch.tensor(shard, dtype=torch.int64, device=device)
            valid = idx >= 0
            safe = idx.clamp_min(0)
            buf = torch.zeros(per, Tx + 3, dtype=torch.int64, device=device)
            buf[:, :Tx] = xs[safe] * valid[:, None]
            buf[:, Tx] = ls[safe] * valid
            buf[:, Tx + 1] = ss[safe] * valid
            buf[:, Tx + 2] = idx
            chunks.append(buf)
    dist.scatter(mine, chunks, src, group=group)
    return (mine[:, :Tx].contiguous(), mine[:, Tx].contiguous(),
            mine[:, Tx + 1].contiguous() if has_sid else None, mine[:, Tx + 2].contiguous())


def gather_audio(audio, n_samples, index_map, total, group=None, dst=0):
    """audio f32[b,1,L_local], n_samples int64[b] valid samples per utterance, index_map from
    scatter_batch.  On `dst` returns a list of `total` 1-D tensors (original order, trimmed to
    their valid length); elsewhere returns None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = audio.device
    L = torch.tensor([audio.shape[2]], dtype=torch.int64, device=dev)
    dist.all_reduce(L, op=dist.ReduceOp.MAX, group=group)
    Lmax = int(L)
    pad = torch.zeros(audio.shape[0], Lmax, dtype=audio.dtype, device=dev)
    pad[:, : audio.shape[2]] = audio[:, 0]
    info = torch.stack([index_map.to(dev), n_samples.to(dev)], dim=1)
    wav_list = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    info_list = [torch.empty_like(info) for _ in range(world)] if rank == dst else None
    dist.gather(pad, wav_list, dst, group=group)
    dist.gather(info, info_list, dst, group=group)
    if rank != dst:
        return None
    out = [None] * total
    for wav, inf in zip(wav_list, info_list):
        for row in range(wav.shape[0]):
            i, n = int(inf[row, 0]), int(inf[row, 1])
            if i >= 0:
                out[i] = wav[row, :n]
    return out


def sharded_infer(net, x, x_lengths, sid, device, group=None, src=0, hop_upsample=256, **infer_kwargs):
    """Scatter -> net.infer on every rank -> gather.  `net` is any object with the
    SynthesizerTrn.infer signature (the CUDA engine in production, a stub in the gloo tests)."""
    rank = dist.get_rank(group)
    total = torch.tensor([0 if x is None else x.shape[0]], dtype=torch.int64, device=device)
    dist.broadcast(total, src, group=group)
    xs, ls, ss, index_map = scatter_batch(x if rank == src else None, x_lengths if rank == src else None,
                                          sid if rank == src else None, device, group, src)
    o, _, y_mask, _ = net.infer(xs, ls.clamp_min(1), ss, **infer_kwargs)
    n_samples = (y_mask.reshape(y_mask.shape[0], -1).sum(dim=1) * hop_upsample).to(torch.int64)
    return gather_audio(o, n_samples, index_map, int(total), group, src)
