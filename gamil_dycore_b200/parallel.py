"""Host-side helpers for the latitude-band decomposition (one process per GPU, torch.distributed for the plumbing).

The data path itself (halo rows, the two-scalar all-reduces) lives inside libgmd over NCCL; this module only mirrors
the band arithmetic of gmd_create (gamil_dycore_b200/csrc/gmd.cu) and gathers bands on the host for output."""
from __future__ import annotations

import numpy as np


def band(rank: int, nranks: int, num_lat: int, polar_band_rows: int = 0):
    """full-latitude rows [r0, r1) owned by `rank` (identical to gmd_get_band): rows split as evenly as possible, the
    first `num_lat % nranks` ranks get one more; with `polar_band_rows` (and nranks >= 3) the first and the last band
    have that many rows and the other ranks share the rest evenly."""
    if nranks >= 3 and polar_band_rows > 0 and 2 * polar_band_rows < num_lat:
        if rank == 0:
            return 0, polar_band_rows
        if rank == nranks - 1:
            return num_lat - polar_band_rows, num_lat
        base, rem = divmod(num_lat - 2 * polar_band_rows, nranks - 2)
        q = rank - 1
        r0 = polar_band_rows + q * base + min(q, rem)
        return r0, r0 + base + (1 if q < rem else 0)
    base, rem = divmod(num_lat, nranks)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def polar_band_rows_for(nranks: int, num_lon: int, num_lat: int, filtered: bool = True) -> int:
    """rows for the two polar bands that balance them against the others (0 = even bands, below 3 ranks).

    A polar band runs the chain sweep -> polar rows -> sweep ... on the rows next to its pole (three links per
    predict_correct, about 51 us + 0.17 us per band row at 3600 columns: the links do not shrink with the band), the other
    bands run one fused kernel per predict_correct (about 14 us + 0.17 us per row).  Measured on B200 at 3600x1801, ms per
    model step: 8 bands, polar bands of 150 / 100 / 66 / 48 / 32 rows: 0.963 / 0.915 / 0.828 / 0.835 / 0.824; 4 bands with
    400-row polar bands 1.33 (profiles/r2_p_*).  The balance point of the two costs, not below 48 rows (the filter rows,
    the plain rows around them and the wide halo must fit) and not above an even share."""
    if nranks < 3 or not filtered:
        return 0
    even = num_lat // nranks
    w = num_lon / 3600.0
    a, b, c0, c1 = 0.172 * w, 0.174 * w, 51.0, 14.0
    mid = nranks - 2
    rp = (b * num_lat / mid + c1 - c0) / (a + 2.0 * b / mid)
    return int(min(even, max(48.0, rp)))


def halo_rows(rank: int, nranks: int, num_lat: int):
    """global rows a rank receives from its neighbours before each stage:
    south: U, V, gd row r0-1; north: U, V row r1 and gd rows r1, r1+1."""
    r0, r1 = band(rank, nranks, num_lat)
    south = [r0 - 1] if rank > 0 else []
    north = [r1, r1 + 1] if rank + 1 < nranks else []
    return south, north


def connect(d, group=None, mode: str = "peer", fallback: bool = True):
    """wire the ranks of one node together.  mode "peer" (default): all-gather the CUDA-IPC blobs and call
    gmd_peer_connect (halo rows and all-reduces over NVLink peer memory, no NCCL on the step path);
    mode "nccl": broadcast an NCCL unique id and call gmd_comm_init (ncclSend/Recv + ncclAllReduce).
    If the peer mapping fails on any rank and `fallback` is set, every rank drops back to "nccl".  Returns the
    mode in use."""
    import torch.distributed as dist
    from . import comm_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    if mode == "peer":
        blobs = [None] * world
        dist.all_gather_object(blobs, d.peer_export(), group=group)
        err = None
        try:
            d.peer_connect(blobs)
        except Exception as e:   # e.g. CUDA IPC not permitted between the ranks' processes
            err = str(e)
        errs = [None] * world
        dist.all_gather_object(errs, err, group=group)   # doubles as the barrier: nobody stores into a neighbour
        if any(errs):                                    # before every rank has mapped it
            if not fallback:
                raise RuntimeError(f"gmd_peer_connect failed: {[e for e in errs if e]}")
            d.peer_disconnect()
            return connect(d, group=group, mode="nccl")
        return "peer"
    elif mode == "nccl":
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=group)
        d.comm_init(uid[0])
        return "nccl"
    else:
        raise ValueError(mode)


def gather_field(local: np.ndarray, num_lat: int, group=None, polar_band_rows: int = 0) -> np.ndarray:
    """`local` is a global-shaped array ([num_lat][num_lon] or, for half-latitude fields, [num_lat-1][num_lon]) in
    which only this rank's band rows are filled (what Dycore.state() returns); all ranks receive the assembled
    field.  Works with any torch.distributed backend."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nrows_global = local.shape[0]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    out = np.zeros_like(local)
    for r in range(world):
        r0, r1 = band(r, world, num_lat, polar_band_rows)
        r1 = min(r1, nrows_global)
        if r1 <= r0:
            continue
        buf = torch.from_numpy(np.ascontiguousarray(local[r0:r1])).to(dev) if r == rank else \
            torch.empty((r1 - r0,) + local.shape[1:], dtype=torch.float64, device=dev)
        dist.broadcast(buf, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        out[r0:r1] = buf.cpu().numpy()
    return out
