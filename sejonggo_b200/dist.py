"""Multi-GPU plumbing: games are independent units, so ranks shard games
with NO collective on the per-simulation path.  NCCL (torch.distributed) is used for the
two real exchanges only, replacing the reference's scp of .h5 models and zipped game
directories (slave_coordinator.py:45-82):
  * broadcast_params  — rank 0's weights to every rank at a model change
  * gather_records    — packed game records to rank 0 at game end (counts, then payload)
"""
import os
import numpy as np
import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise from the torchrun environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_games(n_games, rank, world):
    """Game g lives on rank g mod world; returns this rank's global game ids."""
    return np.arange(rank, n_games, world)


def _flat_items(params):
    for k in sorted(k for k in params if k != 'meta'):
        v = params[k]
        if isinstance(v, dict):
            for kk in sorted(v):
                yield (k, kk), v[kk]
        else:
            yield (k, None), v


def broadcast_params(params, src=0, device=None):
    """In-place broadcast of a model.init_params-style dict (one flat fp32 buffer)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return params
    items = list(_flat_items(params))
    flat = torch.cat([t.reshape(-1).float() for _, t in items])
    if device is not None:
        flat = flat.to(device)
    dist.broadcast(flat, src=src)
    flat = flat.cpu()
    o = 0
    for (k, kk), t in items:
        n = t.numel()
        new = flat[o:o + n].reshape(t.shape).to(t.dtype)
        if kk is None:
            params[k] = new
        else:
            params[k][kk] = new
        o += n
    return params


def gather_records(payload, dst=0, device=None):
    """payload: 1-D uint8 tensor of this rank's packed records.  Two-phase gather: byte counts
    (all_gather), then a gather of the padded payloads to dst.  Returns a list of per-rank tensors on dst, else None."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [payload]
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else payload.device
    count = torch.tensor([payload.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, count)
    mx = int(max(int(c.item()) for c in counts))
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    buf[:payload.numel()] = payload.to(dev)
    bufs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gather_list=bufs, dst=dst)          # payloads travel to dst only (NCCL: grouped send/recv)
    if rank != dst:
        return None
    return [b[:int(c.item())].cpu() for b, c in zip(bufs, counts)]


def pack_records(games):
    """game_data list -> one uint8 tensor (move lists + results; boards are replayable from moves)."""
    out = []
    for g in games:
        mv = np.array([m['move'][0] + 1000 * m['move'][1] for m in g['moves']], np.int32)
        head = np.array([len(mv), -1 if g['winner'] is None else g['winner']], np.int32)
        out.append(head.view(np.uint8))
        out.append(mv.view(np.uint8))
    if not out:
        return torch.zeros(0, dtype=torch.uint8)
    return torch.from_numpy(np.concatenate(out).copy())


def unpack_records(buf):
    a = buf.numpy().view(np.int32)
    games, o = [], 0
    while o < len(a):
        n, w = int(a[o]), int(a[o + 1])
        mv = a[o + 2:o + 2 + n]
        games.append(dict(moves=[(int(m % 1000), int(m // 1000)) for m in mv], winner=None if w < 0 else w))
        o += 2 + n
    return games
