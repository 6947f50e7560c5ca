"""Clip-sharded multi-GPU extraction: one process per GPU, contiguous clip ranges per rank, no
collective on the compute path; ONE all-gather of the rank's (B/G, M, T) mel block when every rank
needs all frames.  The reference has no distributed code (only nn.DataParallel key
stripping, trainer.py:269-272), so this is new surface, not a mirror.

Two gathers:

  mode="nccl"   `dist.all_gather_into_tensor` (NCCL over NVLink / NVSwitch); works on any backend (gloo in the CPU tests).
  mode="fused"  the extraction kernel writes the rank's block straight into its slot of a SYMMETRIC buffer
                (torch.distributed._symmetric_memory: the same allocation on every rank, peer-mapped) and ONE more
                kernel finishes the step (b200mel_gather_pull): it runs the cross-rank barrier itself (system-scope
                release / acquire flags in symmetric memory) and then pulls the peers' blocks with 16-byte loads
                over NVLink — no NCCL call, no host synchronisation, no staging copy of the local block, CUDA-graph
                replayable, and the result IS the symmetric buffer.
                Three buffers rotate: a rank may only overwrite its block of a buffer once every peer has pulled
                it, which its own NEXT barrier implies — so the extraction of step i is ordered after the gather of
                step i-2 (automatic on one stream; with a separate communication stream wait on that gather's
                event, as bench.py does) and still overlaps the gather of step i-1.
                (`ShardedExtractor.__call__` returns a view that stays valid for the next two calls.)
"""
import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced clip range [start, stop) of `rank` (first n % G ranks get one extra clip)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(n_clips, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_gather_mel(local: torch.Tensor, n_clips: Optional[int] = None, group=None) -> torch.Tensor:
    """Gather per-rank (b_r, M, T) blocks into (sum b_r, M, T) on every rank.

    Equal shards use a single `all_gather_into_tensor` (NCCL over NVLink on GPUs; dim-0 contiguous so no
    re-layout).  Ragged shards (n_clips not divisible by the world size) pad the short ranks by one clip
    for the collective and drop the padding afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    b_local = local.shape[0]
    if n_clips is None:
        n_clips = b_local * world
    sizes = [shard_range(n_clips, r, world) for r in range(world)]
    counts = [b - a for a, b in sizes]
    if counts[rank] != b_local:
        raise ValueError(f"rank {rank} holds {b_local} clips, expected {counts[rank]}")
    bmax = max(counts)
    local = local.contiguous()
    if bmax != b_local:
        padded = local.new_zeros((bmax,) + tuple(local.shape[1:]))
        padded[:b_local] = local
        local = padded
    out = local.new_empty((world * bmax,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local, group=group)
    if all(c == bmax for c in counts):
        return out
    return torch.cat([out[r * bmax:r * bmax + counts[r]] for r in range(world)], dim=0)


class SymmetricGather:
    """Peer-memory all-gather of clip-sharded (n_clips, M, T) float32 tensors (mode="fused").

    `slot()` hands out this step's full-size symmetric buffer; the caller writes its own rows [a, b) into it
    (the extraction kernel does, through `out=`), then `finish()` enqueues barrier + pull on the current stream and
    returns the buffer, which now holds every rank's rows.  Buffers are (re)allocated when the shape grows — a
    collective operation, so every rank must call with the same shapes in the same order."""

    N_SLOTS = 3

    def __init__(self, group=None, engine: str = "pull", pull_ctas: int = 0):
        """engine="pull": SM kernel with 16-byte peer loads (to overlap it with an extraction on another stream,
        launch that extraction with reserve_sms=R and give the pull pull_ctas=2 R);
        engine="tma": the same gather driven by the TMA engine (bulk async copies through shared memory, one CTA per
        SM; beside an extraction with reserve_sms=R give it pull_ctas=R — 16 are enough);
        engine="copy": one-CTA barrier kernel + copy-engine transfers (overlaps with an extraction kernel that fills
        the SMs on another stream)."""
        import torch.distributed._symmetric_memory as symm_mem

        if engine not in ("pull", "tma", "copy"):
            raise ValueError("engine must be 'pull', 'tma' or 'copy'")
        self.engine = engine
        self.pull_ctas = int(pull_ctas)  # CTAs of the pull kernel (0 = two per SM); 2 R beside an extraction with reserve_sms=R

        self._symm = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self._bufs: List[torch.Tensor] = []
        self._hdls = []
        self._capacity = 0
        self._step = 0
        self._shape = None
        self._sync = None        # symmetric int32[world + 2]: arrival flags, epoch, ticket (gather_pull_kernel)
        self._sync_hdl = None
        self._call_args = {}     # ctypes argument blocks per (slot, shape): kept alive, reused across steps

    def _ensure(self, numel: int, device: torch.device) -> None:
        if self._sync is None:
            self._sync = self._symm.empty(self.world + 2, dtype=torch.int32, device=device)
            self._sync.zero_()
            self._sync_hdl = self._symm.rendezvous(self._sync, self.group)
            torch.cuda.current_stream(device).synchronize()
            self._sync_hdl.barrier(channel=0)  # every rank's sync words are zero before anyone signals
        if numel <= self._capacity:
            return
        self._call_args = {}
        self._bufs, self._hdls = [], []
        for _ in range(self.N_SLOTS):
            t = self._symm.empty(numel, dtype=torch.float32, device=device)
            self._hdls.append(self._symm.rendezvous(t, self.group))
            self._bufs.append(t)
        self._capacity = numel

    def slot(self, n_clips: int, n_mels: int, n_frames: int, device: torch.device) -> torch.Tensor:
        numel = n_clips * n_mels * n_frames
        self._ensure(numel, device)
        self._shape = (n_clips, n_mels, n_frames)
        return self._bufs[self._step % self.N_SLOTS][:numel].view(n_clips, n_mels, n_frames)

    def finish(self) -> torch.Tensor:
        """Enqueue the gather of the current slot on the current stream (ONE kernel: in-kernel barrier + peer pulls;
        capturable in a CUDA graph) and return the slot, which holds every rank's rows once the kernel has run."""
        from . import _lib

        n_clips, n_mels, n_frames = self._shape
        i = self._step % self.N_SLOTS
        buf, hdl = self._bufs[i], self._hdls[i]
        self._step += 1
        per_clip = n_mels * n_frames
        key = (i, self._shape)
        args = self._call_args.get(key)
        if args is None:
            offs = [shard_range(n_clips, r, self.world)[0] * per_clip for r in range(self.world)] + [n_clips * per_clip]
            args = ((C.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs]),
                    (C.c_void_p * self.world)(*[int(p) for p in self._sync_hdl.buffer_ptrs]),
                    (C.c_int64 * (self.world + 1))(*offs))
            self._call_args[key] = args
        with torch.cuda.device(buf.device):
            st = C.c_void_p(torch.cuda.current_stream(buf.device).cuda_stream)
            if self.engine == "pull":
                rc = _lib.lib().b200mel_gather_pull(buf.data_ptr(), args[0], args[1], self.world, self.rank, args[2],
                                                    self.pull_ctas, st)
            elif self.engine == "tma":
                rc = _lib.lib().b200mel_gather_tma(buf.data_ptr(), args[0], args[1], self.world, self.rank, args[2],
                                                   self.pull_ctas, st)
            else:
                rc = _lib.lib().b200mel_gather_copy(buf.data_ptr(), args[0], args[1], self.world, self.rank, args[2], st)
        _lib.check(rc)
        return buf[:n_clips * per_clip].view(n_clips, n_mels, n_frames)


class ShardedExtractor:
    """Run `module` on this rank's contiguous shard of a clip batch that every rank holds (or can
    index), optionally gathering the result.  `module(wav (b, L)) -> (b, M, T)`.

    mode="nccl": module output, then all_gather_mel.  mode="fused": the module must accept `out=` (the extraction
    modules of this package do); its kernel writes into the symmetric gather buffer and the gather is a device-side
    barrier + one peer-pull kernel (SymmetricGather)."""

    def __init__(self, module, group=None, mode: str = "nccl"):
        if mode not in ("nccl", "fused"):
            raise ValueError("mode must be 'nccl' or 'fused'")
        self.module = module
        self.group = group
        self.mode = mode
        self._sg: Optional[SymmetricGather] = None

    def __call__(self, wav_all: torch.Tensor, gather: bool = True) -> torch.Tensor:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        a, b = shard_range(wav_all.shape[0], rank, world)
        if not (gather and world > 1):
            return self.module(wav_all[a:b])
        if self.mode == "nccl":
            return all_gather_mel(self.module(wav_all[a:b]), wav_all.shape[0], self.group)
        if self._sg is None:
            self._sg = SymmetricGather(self.group)
        n_mels, n_frames = self.out_shape(wav_all.shape[1])
        full = self._sg.slot(wav_all.shape[0], n_mels, n_frames, wav_all.device)
        if b > a:
            self.module(wav_all[a:b], out=full[a:b])
        return self._sg.finish()

    def out_shape(self, n_samples: int) -> Tuple[int, int]:
        """(n_mels, n_frames) of the module's output for clips of n_samples (fused mode needs it before the launch)."""
        m = self.module
        plan = m._plan(torch.device("cuda", torch.cuda.current_device()))
        return plan.cfg.n_mels, plan.out_frames(n_samples)
