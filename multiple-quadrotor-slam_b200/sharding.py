"""
Multi-GPU sharding of the triangulation hot path (one process per GPU, torch.distributed for the plumbing).

Every correspondence is independent in all four solvers (the reference's loop bodies only touch index xi:
Work/python_libs/triangulation_c/triangulation.c:71,110), so the path shards by contiguous point ranges with NO
collective inside the solve.  The only shared inputs are the camera matrices (<= 28 pairs x 2 x 96 B), broadcast once
from rank 0; results are gathered (all_gather over NCCL/NVLink, or gloo in the CPU tests) only when the caller asks
for the assembled map.  The reference has no multi-process code at all; this module is the B200
replacement for its disabled `#pragma omp parallel for` (triangulation_c/setup.py:12-13).
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of rank `rank`: r*ceil(n/G) .. min(n, (r+1)*ceil(n/G))."""
    per = -(-n // world) if world > 0 else n
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def pair_segments(num_cams, total_points, align=256):
    """Segment table of the multi-quadrotor scene: all camera pairs (i < j), equal share of the correspondences (rounded up
    to a multiple of `align` points, the last pair takes the remainder).  Returns a list of (cam_i, cam_j, offset, count)
    covering [0, total_points).  Aligned segment starts keep every warp's 32-point row stores on whole 128-byte lines --
    in local HBM it hardly matters, over NVLink (PeerGather) an unaligned shard costs a third of the link throughput."""
    pairs = [(i, j) for i in range(num_cams) for j in range(i + 1, num_cams)]
    per = -(-total_points // len(pairs))
    if align > 1:
        per = -(-per // align) * align
    segs, off = [], 0
    for (i, j) in pairs:
        cnt = max(0, min(per, total_points - off))
        segs.append((i, j, off, cnt))
        off += cnt
    return segs


def intersect_segments(segs, lo, hi):
    """Pieces of the segment table that fall inside the shard [lo, hi): (cam_i, cam_j, global offset, count)."""
    out = []
    for (i, j, off, cnt) in segs:
        a, b = max(off, lo), min(off + cnt, hi)
        if b > a:
            out.append((i, j, a, b - a))
    return out


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_cameras(cams, src=0, device=None):
    """Broadcast a list/array of 3x4 (or 4x4) camera matrices from `src`; every rank passes an array of the same
    shape (contents ignored on the other ranks).  Works with any initialised backend (NCCL on GPUs, gloo on CPU)."""
    import torch
    dist = _dist()
    arr = np.ascontiguousarray(np.asarray(cams, dtype=np.float64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return arr
    t = torch.from_numpy(arr.copy())
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src)
    return t.cpu().numpy()


def gather_shards(x_shard, n_total, device=None):
    """All-gather equally sized (padded) shards and cut the result back to n_total rows."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(x_shard)[:n_total]
    world = dist.get_world_size()
    per = -(-n_total // world)
    x_shard = np.asarray(x_shard)
    pad = np.zeros((per,) + x_shard.shape[1:], dtype=x_shard.dtype)
    pad[:len(x_shard)] = x_shard
    t = torch.from_numpy(pad.view(np.uint8) if pad.dtype == np.bool_ else pad)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)
    res = out.cpu().numpy()[:n_total]
    return res.view(np.bool_) if x_shard.dtype == np.bool_ else res


class PeerGather:
    """
    The result gather fused into the solver kernels (include/triangl_cuda.h "result mirrors"): the ranks that need the
    map own full-size `x_all (n_total,3)` / `status_all (n_total,)` device arrays, the others map them through CUDA IPC
    (handles are exchanged with `all_gather_object` of the default process group -- NCCL or gloo, plumbing only), and each
    solver call on the local shard stores straight into all of them over NVLink.  After `finish()` the gathered arrays are
    complete.

    gather_dtype : dtype of the gathered x rows.  None = x_dtype.  np.float32 with x_dtype float64 = the map in the
                   reference's SLAM convention (slam2.py:19 sets the output dtype to float32): the rank's own shard stays
                   float64 in `x_local`, the gathered rows travel and are stored as float32 (12 instead of 24 B/point over
                   the links, which bound the gather).
    root         : None = every rank receives the map (all-gather); r = only rank r does (gather): the other ranks
                   allocate no gathered arrays and ship their shard once instead of world-1 times.
    """

    def __init__(self, n_total, x_dtype=np.float64, status_dtype=np.uint8, gather_dtype=None, root=None):
        import triangl_cuda as tc
        dist = _dist()
        self.tc = tc
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > 8:
            raise ValueError("at most 8 ranks")
        self.n_total, self.root = n_total, root
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        self.x_dtype = np.dtype(x_dtype)
        self.gather_dtype = np.dtype(gather_dtype) if gather_dtype is not None else self.x_dtype
        if self.gather_dtype != self.x_dtype and not (self.gather_dtype == np.float32 and self.x_dtype == np.float64):
            raise ValueError("gather_dtype must equal x_dtype, or be float32 for float64 results")
        self.narrow = self.gather_dtype != self.x_dtype
        self.owner = root is None or root == self.rank
        self.x_all = tc.DeviceArray((n_total, 3), self.gather_dtype) if self.owner else None
        self.status_all = tc.DeviceArray((n_total,), status_dtype) if self.owner else None
        n = self.hi - self.lo
        # narrowed gather: the rank's own full-precision result lives apart from the float32 map
        self.x_local = tc.DeviceArray((n, 3), self.x_dtype) if (self.narrow or not self.owner) else None
        self.status_local = tc.DeviceArray((n,), status_dtype) if not self.owner else None
        mine = (tc.ipc_export(self.x_all), tc.ipc_export(self.status_all)) if self.owner else None
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        self.peers = []
        for r, h in enumerate(handles):
            if r != self.rank and h is not None:
                self.peers.append((tc.ipc_import(h[0]), tc.ipc_import(h[1])))
        self.xb = 3 * self.gather_dtype.itemsize
        self.sb = np.dtype(status_dtype).itemsize

    def shard_outputs(self):
        """(x, status) buffers of this rank's shard: pass them as x= / status= of the solver call."""
        n = self.hi - self.lo
        x = self.x_local if self.x_local is not None else self.x_all.view(3 * self.lo, (n, 3))
        st = self.status_local if self.status_local is not None else self.status_all.view(self.lo, (n,))
        return x, st

    def mirror_table(self, sub=0):
        at = self.lo + sub
        table = [(px + at * self.xb, ps + at * self.sb) for (px, ps) in self.peers]
        if self.narrow and self.owner:
            # this rank's own float32 rows are one more "mirror" (in local HBM); its status already lands in status_all
            table.append((self.x_all.ptr + at * self.xb, self.status_all.ptr + at * self.sb))
        return table

    def arm(self, sub=0):
        """Attach the shard addresses inside the gathered arrays to the next device-mode solver call of this thread.
        sub: the call solves the piece of the shard that starts `sub` points into it (one call per camera-pair segment)."""
        self.tc.set_result_mirrors(self.mirror_table(sub), x_f32=self.narrow)

    def egress_bytes_per_point(self):
        """Bytes this rank ships over NVLink per point of its shard."""
        return len(self.peers) * (self.xb + self.sb)

    def finish(self):
        """All ranks' kernels have completed: the gathered arrays are valid everywhere."""
        self.tc.synchronize()
        _dist().barrier()

    def close(self):
        for (px, ps) in self.peers:
            self.tc.ipc_close(px); self.tc.ipc_close(ps)
        self.peers = []


def triangulate_sharded(solve_fn, u1, P1, u2, P2, gather=True, device=None, **kwargs):
    """
    Shard one (u1, P1, u2, P2) batch over the ranks of the default process group.
    `solve_fn` has the reference signature (e.g. triangulation.linear_LS_triangulation); every rank passes the same
    full-size u1/u2 (or at least its own range filled in).  Returns (x, status) assembled on every rank when `gather`
    is true, else this rank's shard and its (lo, hi).
    """
    dist = _dist()
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if on else 0
    world = dist.get_world_size() if on else 1
    n = len(u1)
    cams = broadcast_cameras(np.stack([np.asarray(P1, dtype=np.float64)[0:3], np.asarray(P2, dtype=np.float64)[0:3]]),
                             0, device)
    lo, hi = shard_range(n, rank, world)
    x, status = solve_fn(u1[lo:hi], cams[0], u2[lo:hi], cams[1], **kwargs)
    if not gather:
        return x, status, (lo, hi)
    return gather_shards(x, n, device), gather_shards(status, n, device)


def triangulate_pairs_sharded(solve_fn, u_by_cam, cams, segs, gather=True, device=None, **kwargs):
    """
    Multi-camera scene: `segs` from pair_segments(); u_by_cam[(i, j)] = (u_i, u_j) arrays of that pair's matches
    (length = count).  The concatenated correspondence array is sharded by range, every rank holds all camera matrices.
    """
    dist = _dist()
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if on else 0
    world = dist.get_world_size() if on else 1
    cams = broadcast_cameras(np.stack([np.asarray(P, dtype=np.float64)[0:3] for P in cams]), 0, device)
    total = sum(c for (_, _, _, c) in segs)
    lo, hi = shard_range(total, rank, world)
    xs, sts = [], []
    seg_off = {(i, j): off for (i, j, off, _) in segs}
    for (i, j, off, cnt) in intersect_segments(segs, lo, hi):
        a = off - seg_off[(i, j)]
        ui, uj = u_by_cam[(i, j)]
        x, st = solve_fn(ui[a:a + cnt], cams[i], uj[a:a + cnt], cams[j], **kwargs)
        xs.append(np.asarray(x)); sts.append(np.asarray(st))
    x = np.concatenate(xs) if xs else np.zeros((0, 3))
    st = np.concatenate(sts) if sts else np.zeros((0,), dtype=np.bool_)
    if not gather:
        return x, st, (lo, hi)
    return gather_shards(x, total, device), gather_shards(st, total, device)


def triangulate_multiview_sharded(solve_fn, us, Ps, valid=None, gather=True, device=None, **kwargs):
    """
    m-view scene: shard the POINT axis of `us` (m, N, 2) / `valid` (m, N) over the ranks; every
    rank receives all m camera matrices from rank 0.  `solve_fn` has the signature of
    triangulation.multiview_LS_triangulation(us, Ps, valid, ...).  Returns (x, status) assembled on every rank when
    `gather` is true, else this rank's shard and its (lo, hi).  No data-path collective besides the optional gather.
    """
    dist = _dist()
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if on else 0
    world = dist.get_world_size() if on else 1
    n = us.shape[1]
    cams = broadcast_cameras(np.stack([np.asarray(P, dtype=np.float64)[0:3] for P in Ps]), 0, device)
    lo, hi = shard_range(n, rank, world)
    x, status = solve_fn(us[:, lo:hi], list(cams), None if valid is None else valid[:, lo:hi], **kwargs)
    if not gather:
        return x, status, (lo, hi)
    return gather_shards(x, n, device), gather_shards(status, n, device)

