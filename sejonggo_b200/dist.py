"""Multi-GPU plumbing: games are independent units, so ranks shard games
with NO collective on the per-simulation path.  NCCL (torch.distributed) is used for the
two real exchanges only, replacing the reference's scp of .h5 models and zipped game
directories (slave_coordinator.py:45-82):
  * broadcast_model — rank 0's network to every rank at a model change: ONE blob of folded weights (tower convs in
                      bf16, 48 MB at 20 blocks), device to device
  * gather_rows     — the packed game records (records.py rows: board, move, value, policy_target per ply) to rank 0:
                      row counts by all_gather, then the payloads device to device
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


def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


HEADER_BYTES = 256


def broadcast_model(model, src=0, device=None, max_positions=None):
    """Every rank returns a TowerModel with rank `src`'s weights and name.  `model` is only read on `src` (others may
    pass None).  Two collectives: a 256-byte header (name, board size, blocks), then the weight blob, which is
    packed on the device on `src` and consumed in place on the others — no host copy on either side."""
    from . import model as M
    if _world() == 1:
        return model
    rank = dist.get_rank()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    head = torch.zeros(HEADER_BYTES, dtype=torch.uint8, device=dev)
    if rank == src:
        nm = model.name.encode("utf-8")[:HEADER_BYTES - 16]
        h = np.zeros(HEADER_BYTES, np.uint8)
        h[:4] = np.array([model.size], np.int32).view(np.uint8)
        h[4:8] = np.array([model.n_blocks], np.int32).view(np.uint8)
        h[8:12] = np.array([len(nm)], np.int32).view(np.uint8)
        h[16:16 + len(nm)] = np.frombuffer(nm, np.uint8)
        head.copy_(torch.from_numpy(h))
    dist.broadcast(head, src=src)
    h = head.cpu().numpy()
    size, n_blocks, ln = (int(x) for x in h[:12].view(np.int32))
    name = bytes(h[16:16 + ln]).decode("utf-8")
    _, total = M.blob_layout(size, n_blocks)
    blob = model.blob(dev) if rank == src else torch.empty(total, dtype=torch.uint8, device=dev)
    dist.broadcast(blob, src=src)
    if rank == src:
        return model
    mp = max_positions or (model.max_positions if model is not None else 8192)
    return M.TowerModel(name, folded=(M.unpack_blob(blob, size, n_blocks), size, n_blocks), max_positions=mp)


def blob_bytes(model):
    from . import model as M
    return M.blob_layout(model.size, model.n_blocks)[1]


def gather_rows(rows, dst=0):
    """rows: int32 [n][RW] tensor of this rank's record rows (on the device the process group communicates on).
    Returns on `dst` the list of every rank's rows (tensors on that device), elsewhere None.  Row counts travel by
    all_gather; each payload by ONE exact-size send/recv (NCCL: device to device over NVLink)."""
    if _world() == 1:
        return [rows]
    world, rank = dist.get_world_size(), dist.get_rank()
    RW = int(rows.shape[1])
    count = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros(1, dtype=torch.int64, device=rows.device) for _ in range(world)]
    dist.all_gather(counts, count)
    counts = [int(c.item()) for c in counts]
    if rank == dst:
        out, ops = [], []
        for r in range(world):
            if r == dst:
                out.append(rows)
                continue
            buf = torch.empty((counts[r], RW), dtype=torch.int32, device=rows.device)
            out.append(buf)
            if counts[r]:
                ops.append(dist.P2POp(dist.irecv, buf, r))
        for q in (dist.batch_isend_irecv(ops) if ops else []):
            q.wait()
        return out
    if rows.shape[0]:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, rows.contiguous(), dst)]):
            q.wait()
    return None


def gather_per_rank(values, device=None):
    """Small per-rank diagnostics (a list of floats) from every rank: -> float64 [world][len(values)] on all ranks."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if _world() == 1:
        return t.reshape(1, -1).cpu().numpy()
    outs = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return torch.stack(outs).cpu().numpy()
