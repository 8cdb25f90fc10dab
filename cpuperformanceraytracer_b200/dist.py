"""dist.py -- multi-GPU sharding of the path loop (one process per GPU, torch.distributed for the
plumbing; NCCL over NVLink on GPUs, gloo in the CPU tests of the host logic).

The reference has no distributed code.  Its path loop shards trivially: every (pixel, frame)
sample re-seeds its RNG from (x, y, iFrame) (demofox_path_tracing_optimization_v4.cpp:1096-1101),
so frames are independent streams and only the accumulation couples them.

  spp-shard  rank r renders a contiguous block of the 1-based frame range into a SUM buffer
             (B200PT_ACCUM_SUM); one all-reduce(sum) of the W*H*3 f32 buffer; then every rank
             scales by 1/(N+1) (b200pt_finalize_sum) -- the value the reference's running average
             reaches after N calls on a zeroed buffer.  Same samples, different summation order:
             agrees with the sequential render to ~sqrt(frames) x 2^-24 relative (6e-7 at 16 frames, 6e-6 at 1024), not bit for bit.
  tile-shard rank r renders tile rows [a, b) of every frame; the buffer is band-major (a row of
             tiles is contiguous, RenderTile v4.cpp:1189-1194) so each rank owns one contiguous
             span and an all-gather reassembles the image with no repacking.  Bit-identical to the
             single-GPU render.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class FrameShard:
    first_frame: int  # 1-based iFrame of the first render call of this rank
    nframes: int


def shard_frames(total_frames: int, world_size: int, rank: int, first_frame: int = 1) -> FrameShard:
    """Contiguous, near-equal blocks of the frame range [first_frame, first_frame + total_frames)."""
    if world_size <= 0 or not (0 <= rank < world_size) or total_frames < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total_frames, world_size)
    n = base + (1 if rank < rem else 0)
    start = first_frame + rank * base + min(rank, rem)
    return FrameShard(start, n)


@dataclass(frozen=True)
class TileRowShard:
    first_tile_row: int
    num_tile_rows: int
    float_offset: int  # offset of the rank's span in the accumulation buffer
    float_count: int


def shard_tile_rows(width: int, height: int, num_tiles_y: int, world_size: int, rank: int, row_costs=None) -> TileRowShard:
    """Tile-row bands: one contiguous span of the band-major buffer per rank.  row_costs (one weight per tile row)
    makes the bands near-equal in COST instead of near-equal in rows: sky rows are ~10x cheaper than scene rows."""
    if height % num_tiles_y or world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad shard request")
    band = (height // num_tiles_y) * width * 3
    if row_costs is None:
        base, rem = divmod(num_tiles_y, world_size)
        n = base + (1 if rank < rem else 0)
        start = rank * base + min(rank, rem)
        return TileRowShard(start, n, start * band, n * band)
    if len(row_costs) != num_tiles_y or min(row_costs) < 0:
        raise ValueError("one non-negative cost per tile row")
    total = float(sum(row_costs))
    t, acc, bounds = 0, 0.0, [0]
    for r in range(world_size):
        target = total * (r + 1) / world_size
        while t < num_tiles_y and (r == world_size - 1 or acc + 0.5 * row_costs[t] <= target):
            acc += row_costs[t]
            t += 1
        bounds.append(t)
    start, n = bounds[rank], bounds[rank + 1] - bounds[rank]
    return TileRowShard(start, n, start * band, n * band)


def tile_row_costs(cull_rects, width: int, height: int, num_tiles_y: int, traced_weight: float = 10.0):
    """Relative cost of every tile row from the camera-culling rectangles (api.cull_rects): a pixel whose jitter footprint
    touches no rectangle never traces the scene (weight 1), the others do (weight traced_weight)."""
    import numpy as np
    th = height // num_tiles_y
    if cull_rects is None:
        return [1.0] * num_tiles_y
    xs = np.arange(0, width, 4, dtype=np.float32)[None, :]
    costs = []
    for ty in range(num_tiles_y):
        ys = (height - 1 - np.arange(ty * th, (ty + 1) * th, 4, dtype=np.float32))[:, None]  # flipped rows, like fragCoord.y
        hit = np.zeros((ys.shape[0], xs.shape[1]), dtype=bool)
        for (x0, y0, x1, y1) in np.asarray(cull_rects, dtype=np.float32).reshape(-1, 4):
            hit |= (xs + 0.5 >= x0) & (xs - 0.5 <= x1) & (ys + 0.5 >= y0) & (ys - 0.5 <= y1)
        costs.append(float((~hit).sum() + traced_weight * hit.sum()))
    return costs


def finalize_scale(last_frame: int) -> float:
    """1/(N+1) for a job whose last render call is frame N: the reference's running average on a zeroed buffer is
    biased by design (iFrame starts at 1 and the blend factor is 1/(iFrame+1))."""
    return 1.0 / (float(last_frame) + 1.0)


def reduce_sum_(tensor, group=None):
    """All-reduce(sum) of a rank's SUM buffer in place (NCCL on CUDA tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


class SppShardedRenderer:
    """spp-sharded render of one image on this rank's GPU.  The accumulation buffer is a torch CUDA
    tensor (so NCCL can reduce it in place) bound to the engine through b200pt_bind_device_target;
    kernels run on torch's current stream."""

    def __init__(self, renderer_factory, width, height, ntx, nty, rank, world_size, device):
        import torch
        from . import api
        self.torch = torch
        self.rank, self.world = rank, world_size
        self.width, self.height, self.nty = width, height, nty
        self.comm_stream = None
        self.device = torch.device("cuda", device)
        self.r = renderer_factory(accum_mode=api.ACCUM_SUM, device=device)
        self.r.resize(width, height, ntx, nty)
        self.buf = torch.zeros(width * height * 3, dtype=torch.float32, device=self.device)
        self.r.bind_device_target(self.buf.data_ptr())
        # a dedicated (non-default) stream shared by the kernels, the memset and the collective
        self.stream = torch.cuda.Stream(self.device)
        self.r.set_stream(self.stream.cuda_stream)
        torch.cuda.synchronize(self.device)

    def render(self, total_frames, first_frame=1, resume=False, bands=1):
        """Renders frames [first_frame, first_frame + total_frames) of the job, sharded over the ranks, reduces, scales.
        Asynchronous.  resume=False: a fresh job (first_frame must be 1): every rank zeroes its SUM buffer.
        resume=True: rank 0's buffer holds the running average after first_frame - 1 render calls (any caller state
        for first_frame = 1); it is turned back into a sum, A * first_frame (the reference's blend factor is
        1/(iFrame + 1)), the other ranks start from zero, and the result is the reference's average
        after first_frame - 1 + total_frames calls.
        bands > 1: the image is rendered in `bands` groups of tile rows; a band is one contiguous span of the
        tile-major buffer, and its all-reduce + scale run on a second stream while the next band renders, so the
        exchange of a large image (8192 x 8192: 805 MB) hides behind the render."""
        if first_frame < 1 or (first_frame != 1 and not resume):
            raise ValueError("a fresh job starts at frame 1; pass resume=True to continue from rank 0's buffer")
        torch = self.torch
        sh = shard_frames(total_frames, self.world, self.rank, first_frame)
        last = first_frame - 1 + total_frames
        bands = max(1, min(int(bands), self.nty))
        with torch.cuda.stream(self.stream):
            if resume and self.rank == 0:
                if first_frame > 1:
                    self.r.scale_target(float(first_frame))
            else:
                self.buf.zero_()
        if bands == 1:
            with torch.cuda.stream(self.stream):
                self.r.frame_counter = sh.first_frame - 1
                self.r.render_frames(sh.nframes, sync=False)
                reduce_sum_(self.buf)
                self.r.finalize_sum(last)
        else:
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream(self.device)
            per_row = self.width * (self.height // self.nty) * 3
            scale = finalize_scale(last)
            for b in range(bands):
                row0, row1 = self.nty * b // bands, self.nty * (b + 1) // bands
                with torch.cuda.stream(self.stream):
                    self.r.set_tile_row_range(row0, row1 - row0)
                    self.r.frame_counter = sh.first_frame - 1
                    self.r.render_frames(sh.nframes, sync=False)
                    done = torch.cuda.Event()
                    done.record(self.stream)
                self.comm_stream.wait_event(done)
                with torch.cuda.stream(self.comm_stream):
                    off, cnt = row0 * per_row, (row1 - row0) * per_row
                    reduce_sum_(self.buf[off:off + cnt])
                    self.r.scale_target_span(off, cnt, scale, self.comm_stream.cuda_stream)
            self.r.set_tile_row_range(0, 0)  # back to the whole image
            self.stream.wait_stream(self.comm_stream)
        self.r.frame_counter = last
        return self.buf

    def render_host_slices(self, host_state, total_frames):
        """The reference-facing call for N ranks whose caller keeps its accumulation buffer in host memory that EVERY rank
        can address (POSIX shared memory, page-locked in each process): rank r moves only floats [lo_r, hi_r) of it, over
        its own PCIe link, in both directions.  The caller's state enters the sum exactly once (each slice is uploaded by
        exactly one rank, the rest of every SUM buffer starts at zero); after the all-reduce every rank holds the finished
        image and writes its slice back.  host_state: a CPU float32 tensor of W*H*3 elements, the same memory on all
        ranks.  Blocks until this rank's slice is in host memory."""
        torch = self.torch
        n = self.buf.numel()
        base, rem = divmod(n // 4, self.world)  # float4-aligned slices
        lo = 4 * (self.rank * base + min(self.rank, rem))
        hi = lo + 4 * (base + (1 if self.rank < rem else 0))
        sh = shard_frames(total_frames, self.world, self.rank, 1)
        with torch.cuda.stream(self.stream):
            self.buf.zero_()
            if hi > lo:
                self.buf[lo:hi].copy_(host_state[lo:hi], non_blocking=True)
            self.r.frame_counter = sh.first_frame - 1
            self.r.render_frames(sh.nframes, sync=False)
            reduce_sum_(self.buf)
            self.r.finalize_sum(total_frames)
            if hi > lo:
                host_state[lo:hi].copy_(self.buf[lo:hi], non_blocking=True)
        self.r.frame_counter = total_frames
        self.stream.synchronize()
        return lo, hi

    def close(self):
        self.r.close()


class TileShardedRenderer:
    """tile-shard render: rank r renders tile rows [a, b) of every frame with the reference's running
    average (bit-identical to a single-GPU render); the contiguous spans are then exchanged so that
    every rank holds the whole image (one broadcast per rank span; spans may differ in size)."""

    def __init__(self, renderer_factory, width, height, ntx, nty, rank, world_size, device):
        import torch
        from . import api
        self.torch = torch
        self.rank, self.world = rank, world_size
        self.device = torch.device("cuda", device)
        self.r = renderer_factory(accum_mode=api.ACCUM_RUNNING_AVERAGE, device=device)
        costs = None
        if not self.r.params.disable_camera_culling:
            costs = tile_row_costs(api.cull_rects(self.r.params.profile, width, height), width, height, nty)
        self.shards = [shard_tile_rows(width, height, nty, world_size, r, costs) for r in range(world_size)]
        self.r.resize(width, height, ntx, nty)
        self.buf = torch.zeros(width * height * 3, dtype=torch.float32, device=self.device)
        self.r.bind_device_target(self.buf.data_ptr())
        me = self.shards[rank]
        self.r.set_tile_row_range(me.first_tile_row, me.num_tile_rows)
        self.stream = torch.cuda.Stream(self.device)
        self.r.set_stream(self.stream.cuda_stream)
        torch.cuda.synchronize(self.device)

    def render(self, nframes, gather=True):
        """nframes more render calls on this rank's rows (frame counter carries on), then the exchange."""
        import torch.distributed as dist
        with self.torch.cuda.stream(self.stream):
            if self.shards[self.rank].num_tile_rows > 0:
                self.r.render_frames(nframes, sync=False)
            else:
                self.r.frame_counter = self.r.frame_counter + nframes
            if gather and self.world > 1:
                for src, sh in enumerate(self.shards):
                    if sh.float_count:
                        dist.broadcast(self.buf[sh.float_offset:sh.float_offset + sh.float_count], src=src)
        return self.buf

    def close(self):
        self.r.close()
