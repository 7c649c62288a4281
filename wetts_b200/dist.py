"""Batch-sharded multi-GPU inference: one process per GPU, weights replicated,
utterances dealt to ranks, no collective inside the model.  The only communication is the
"trivial batch scatter/gather" the north_star names: rank 0 holds the request batch, scatters
ids / lengths / speaker ids (KBs), every rank synthesises its shard, and the waveforms are
gathered back to rank 0 (NCCL over NVLink when the tensors are CUDA tensors; the same code runs
on gloo/CPU tensors for the host-logic tests).

The reference has no inference-time parallelism at all (inference.py:82-100 is batch 1 on one
device); its only sharding logic is the training-time DistributedBucketSampler
(wetts/vits/data_utils.py:228-346), whose length-balancing idea plan_shards() follows.

Parity caveat: the reference's implicit RNG draw depends on the batch composition,
so sharded == unsharded holds bit for bit only with explicit per-utterance noise; `noise_w`,
`noise_z` and `durations` given on the source rank are therefore scattered row by row with the ids.
One exception remains, and the reference model has it too: HiFi-GAN does not mask, so the last samples
of an utterance depend on whether batch padding or the end of the tensor follows it.  The longest
utterance of a shard that was not the longest of the whole batch can therefore differ inside the
vocoder's receptive field of its end (< 16 frames; z and z_p are identical; measured on two B200s,
tests/test_dist_gpu.py, tools/diag_batch_invariance.py).
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


def scatter_rows(t, plan, row_shape, dtype, device, group=None, src=0):
    """Scatter the rows of a per-utterance tensor `t` ([B, ...] on `src`, None elsewhere) following `plan`
    (known on `src` only).  Padding slots receive zeros.  Returns this rank's [per, ...] tensor."""
    rank = dist.get_rank(group)
    per = row_shape[0]
    mine = torch.empty(row_shape, dtype=dtype, device=device)
    chunks = None
    if rank == src:
        td = t.to(device=device, dtype=dtype)
        chunks = []
        for shard in plan:
            idx = torch.tensor(shard, dtype=torch.int64, device=device)
            rows = td[idx.clamp_min(0)]
            rows = rows * (idx >= 0).to(dtype).reshape([per] + [1] * (rows.dim() - 1))
            chunks.append(rows.contiguous())
    dist.scatter(mine, chunks, src, group=group)
    return mine


def scatter_batch(x, x_lengths, sid, device, group=None, src=0, return_plan=False):
    """Rank `src` passes the full batch (x int64[B,Tx], x_lengths int64[B], sid int64[B] or None);
    other ranks pass None.  Returns this rank's (x, x_lengths, sid, index_map): index_map[i] is
    the position of local utterance i in the original batch (-1 = padding slot, length 0 -> the
    engine still synthesises >= 1 frame for it; gather_audio drops it)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    plan = None
    if rank == src:
        plan = plan_shards(x_lengths.tolist(), world)
        meta[0], meta[1], meta[2] = len(plan[0]), x.shape[1], 0 if sid is None else 1
    dist.broadcast(meta, src, group=group)
    per, Tx, has_sid = (int(v) for v in meta.tolist())
    # one packed int64 buffer per rank: [per, Tx + 3] = ids | length | sid | original index
    packed = None
    if rank == src:
        ls = x_lengths.to(torch.int64)
        ss = sid.to(torch.int64) if has_sid else torch.zeros_like(ls)
        packed = torch.cat([x.to(torch.int64), ls[:, None], ss[:, None],
                            torch.arange(x.shape[0], dtype=torch.int64, device=x.device)[:, None]], dim=1)
    mine = scatter_rows(packed, plan, (per, Tx + 3), torch.int64, device, group, src)
    if rank == src:
        # padding slots must carry index -1 (scatter_rows zeroes them)
        pass
    idx_map = mine[:, Tx + 2].clone()
    idx_map[mine[:, Tx] == 0] = -1          # a real utterance has length >= 1; zeroed rows are padding slots
    out = (mine[:, :Tx].contiguous(), mine[:, Tx].contiguous(),
           mine[:, Tx + 1].contiguous() if has_sid else None, idx_map)
    return out + (plan,) if return_plan else out


def gather_audio(audio, n_samples, index_map, total, group=None, dst=0, as_list=True):
    """audio f32[b,1,L_local], n_samples int64[b] valid samples per utterance, index_map from
    scatter_batch.  On `dst` returns a list of `total` 1-D tensors (original order, trimmed to
    their valid length) -- or, with as_list=False, the raw (wav [world*b, Lmax], info int64[world*b, 2] on the
    host: original index, valid samples) pair; elsewhere returns None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = audio.device
    L = torch.tensor([audio.shape[2]], dtype=torch.int64, device=dev)
    dist.all_reduce(L, op=dist.ReduceOp.MAX, group=group)
    Lmax = int(L)
    b = audio.shape[0]
    if audio.shape[2] == Lmax:
        pad = audio[:, 0].contiguous()
    else:
        pad = torch.zeros(b, Lmax, dtype=audio.dtype, device=dev)
        pad[:, : audio.shape[2]] = audio[:, 0]
    info = torch.stack([index_map.to(dev), n_samples.to(dev)], dim=1).contiguous()
    wav_all = info_all = None
    wav_list = info_list = None
    if rank == dst:
        wav_all = torch.empty(world * b, Lmax, dtype=audio.dtype, device=dev)
        info_all = torch.empty(world * b, 2, dtype=torch.int64, device=dev)
        wav_list = list(wav_all.split(b, dim=0))       # views: the gather lands in one contiguous buffer
        info_list = list(info_all.split(b, dim=0))
    dist.gather(pad, wav_list, dst, group=group)
    dist.gather(info, info_list, dst, group=group)
    if rank != dst:
        return None
    info_host = info_all.cpu()                          # ONE device->host sync for the whole table
    if not as_list:
        return wav_all, info_host
    out = [None] * total
    for row, (i, n) in enumerate(info_host.tolist()):
        if i >= 0:
            out[i] = wav_all[row, :n]
    return out


def sharded_infer(net, x, x_lengths, sid, device, group=None, src=0, hop_upsample=256, as_list=True,
                  noise_w=None, noise_z=None, durations=None, **infer_kwargs):
    """Scatter -> net.infer on every rank -> gather.  `net` is any object with the
    SynthesizerTrn.infer signature (the CUDA engine in production, a stub in the gloo tests).
    Optional per-utterance tensors on `src` (noise_w [B,2,Tx], noise_z [B,C,Tmax], durations [B,1,Tx]) are
    scattered with the ids so that sharded == unsharded bit for bit (up to the padding effect at the end of a shard's
    longest utterance described in the module docstring)."""
    rank = dist.get_rank(group)
    meta = torch.zeros(8, dtype=torch.int64, device=device)
    if rank == src:
        meta[0] = x.shape[0]
        for j, t in enumerate((noise_w, noise_z, durations)):
            if t is not None:
                meta[1 + 2 * j], meta[2 + 2 * j] = t.shape[1], t.shape[2]
    dist.broadcast(meta, src, group=group)
    m = meta.tolist()
    total = int(m[0])
    xs, ls, ss, index_map, plan = scatter_batch(x if rank == src else None, x_lengths if rank == src else None,
                                                sid if rank == src else None, device, group, src, return_plan=True)
    per = xs.shape[0]
    extra = {}
    for j, (name, t) in enumerate((("noise_w", noise_w), ("noise_z", noise_z), ("durations", durations))):
        c, tlen = int(m[1 + 2 * j]), int(m[2 + 2 * j])
        if c > 0:
            extra[name] = scatter_rows(t if rank == src else None, plan, (per, c, tlen), torch.float32, device, group, src)
    o, _, y_mask, _ = net.infer(xs, ls.clamp_min(1), ss, **extra, **infer_kwargs)
    n_samples = (y_mask.reshape(y_mask.shape[0], -1).sum(dim=1) * hop_upsample).to(torch.int64)
    return gather_audio(o, n_samples, index_map, total, group, src, as_list=as_list)
