"""torchrun worker for tests/test_gpu_multi.py: every rank extracts its contiguous clip shard on its own GPU and
gathers the mel frames; the gathered tensor must be BIT-EQUAL to the single-GPU extraction of the whole batch,
for even and ragged shards and for every gather mode (NCCL all-gather / fused peer-memory epilogue).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/dist_gather_check.py
Prints one line `DIST_GATHER_OK <world>` on rank 0 when every check passed."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mel_oracle as mo  # noqa: E402  (test infrastructure: input synthesis + checker)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from pytorch_sound_b200.distributed import ShardedExtractor, shard_range
    from pytorch_sound_b200.models.transforms import LogMelSpectrogram

    geo = dict(sample_rate=22050, mel_size=80, n_fft=1024, win_length=1024, hop_length=256, min_db=-50, max_db=30,
               mel_min=0.0, mel_max=8000.0)
    lm = LogMelSpectrogram(**geo).to(dev)
    ok = True
    for n_clips, L in [(64, 22050), (8 * world, 5000), (8 * world + 3, 7001), (world + 1, 2000)]:
        x = torch.from_numpy(mo.synth_clips(n_clips, L, 22050, seed=99 + n_clips)).to(dev)
        single = lm(x)                                   # the whole batch on this GPU
        for mode in ("nccl", "fused", "fused-tma"):
            ext = ShardedExtractor(lm, mode=mode.split("-")[0])
            if mode == "fused-tma":   # the same fused gather driven by the TMA engine (b200mel_gather_tma)
                from pytorch_sound_b200.distributed import SymmetricGather
                ext._sg = SymmetricGather(engine="tma", pull_ctas=8)
            for rep in range(2):                         # twice: buffers / barriers are reused across steps
                got = ext(x)
                a, b = shard_range(n_clips, rank, world)
                same = got.shape == single.shape and torch.equal(got, single)
                local_only = ext(x, gather=False)
                same = same and torch.equal(local_only, single[a:b])
                if not same:
                    print(f"rank {rank}: MISMATCH n_clips={n_clips} L={L} mode={mode} rep={rep}", flush=True)
                ok = ok and same
        # and against the float64 oracle on rank 0's view
        ref = mo.log_mel_spectrogram(x.cpu().numpy(), **geo)
        ok = ok and mo.parity_error(single.cpu().numpy(), ref) < 1e-4
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print(f"DIST_GATHER_OK {world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
