#!/usr/bin/env python
"""Times prebuilt variants of the library at the C2 shape the way bench.py times `value` (CUDA graph of 8 launches over
8 distinct batches, replayed, CUDA events) — variants are built HERE (no GPU needed) so the GPU box only runs them:

    python tools/tc_bench.py build  name1="-DFOO" name2="-DBAR ENV:B200MEL_TC=1"     # -> build/variants/var_<name>.so
    python tools/tc_bench.py run [C2|C3]                                            # on the GPU box: every var_*.so
    python tools/tc_bench.py timing                                                  # phase accounting of var_timing.so

ENV:K=V words become environment variables of the timed child instead of nvcc flags (stored next to the .so)."""
import ctypes as C
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "pytorch_sound_b200")
VAR = os.path.join(ROOT, "build", "variants")  # git-ignored, but travels to the GPU box

CHILD = r'''
import sys, torch
sys.path.insert(0, %(root)r)
from pytorch_sound_b200 import _lib
_lib.LIB_PATH = %(lib)r
from pytorch_sound_b200.models.transforms import LogMelSpectrogram
B, L = %(B)d, %(L)d
m = LogMelSpectrogram(*%(geo)r).cuda()
g0 = torch.Generator(device="cuda").manual_seed(1)
n = %(n)d
xs = [torch.randn(B, L, device="cuda", generator=g0) * 0.1 for _ in range(n)]
for x in xs: m(x)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        outs = [m(x) for x in xs]
torch.cuda.synchronize()
for _ in range(5): g.replay()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / (50 * n) * 1e3)
torch.save(outs[0].cpu(), %(out)r)
print("%%.2f" %% best)
'''


def build_variants(specs):
    from pytorch_sound_b200 import build
    procs = []
    for spec in specs:
        name, _, flags = spec.partition("=")
        words = [w for w in flags.split() if not w.startswith("ENV:")]
        env = {w[4:].partition("=")[0]: w[4:].partition("=")[2] for w in flags.split() if w.startswith("ENV:")}
        os.makedirs(VAR, exist_ok=True)
        lib = os.path.join(VAR, f"var_{name}.so")
        with open(lib + ".json", "w") as f:
            json.dump({"flags": flags, "env": env}, f)
        cmd = [build.find_nvcc()] + build.NVCC_FLAGS + words + ["-o", lib, "b200mel.cu"]
        procs.append((name, subprocess.Popen(cmd, cwd=build.CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, pr in procs:
        out, _ = pr.communicate()
        print(f"[{name}] {'ok' if pr.returncode == 0 else 'BUILD FAILED'}\n{out[-1500:] if pr.returncode else ''}")


def run_variants(workload):
    import torch
    B, L, n, geo = {"C2": (256, 22050, 8, (22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.)),
                    "C3": (256, 88200, 3, (22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.)),
                    "C4": (16, 441000, 8, (44100, 128, 2048, 2048, 512, -50, 30, 0., None)),
                    "C5": (8192, 8000, 2, (16000, 80, 1024, 1024, 256, -50, 30, 0., 8000.))}[workload]
    ref = None
    libs = [os.path.join(PKG, "libb200mel.so")] + sorted(glob.glob(os.path.join(VAR, "var_*.so")))
    for lib in libs:
        meta = json.load(open(lib + ".json")) if os.path.exists(lib + ".json") else {"flags": "shipped", "env": {}}
        if "STC_TIMING" in meta["flags"]:
            continue
        env = dict(os.environ)
        env.update(meta["env"])
        out = os.path.join(ROOT, "gpurun_out", "var_out.pt")
        r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, lib=lib, out=out, B=B, L=L, n=n, geo=geo)],
                           capture_output=True, text=True, env=env)
        tag = os.path.basename(lib)
        if r.returncode:
            print(f"[{tag}: {meta['flags']}] RUN FAILED\n{r.stderr[-1500:]}", flush=True)
            continue
        y = torch.load(out)
        if ref is None:
            ref = y
        print(f"[{tag}: {meta['flags']}] {workload} {r.stdout.strip()} us/launch   max|y - y_shipped| = {(y - ref).abs().max().item():.3g}", flush=True)


def timing():
    import torch
    from pytorch_sound_b200 import _lib
    lib_path = os.path.join(VAR, "var_timing.so")
    _lib.LIB_PATH = lib_path
    lib = _lib.lib()
    lib.b200mel_debug_set_buffer.argtypes = [C.c_void_p]
    lib.b200mel_debug_set_tc_mode.argtypes = [C.c_int]
    lib.b200mel_debug_set_tc_mode(1)
    from pytorch_sound_b200.models.transforms import LogMelSpectrogram
    m = LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.).cuda()
    names = [["poll / idle", "TMA issue", "M2 issue (8 MMAs + commit)", "M1 issue (8 MMAs + commit)", "", "", "", ""],
             ["loop", "wait D1 full (M1 done)", "wait A2 free (M2(j-2) done)", "P2 compute + stores", "", "", "", ""],
             ["loop", "wait D2 full (M2 done)", "wait mag tile free", "P3 compute", "", "", "", ""],
             ["loop", "wait stage (TMA)", "loads + max", "edge barrier", "scale + wait S free", "convert + stores", "", ""],
             ["loop + geometry", "wait mag full", "P4 (mel, epilogue, stores)", "", "", "", "", ""]]
    roles = ["issuer", "twiddle warp 0", "spectrum warp 8", "edge warp 12", "mel warp 16"]
    for B, L in ((256, 22050), (256, 88200)):
        x = torch.randn(B, L, device="cuda") * 0.1
        m(x)
        torch.cuda.synchronize()
        dbg = torch.zeros(48, dtype=torch.int64, device="cuda")
        lib.b200mel_debug_set_buffer(dbg.data_ptr())
        m(x)
        torch.cuda.synchronize()
        lib.b200mel_debug_set_buffer(None)
        d = dbg.cpu().tolist()
        T = L // 256 + 1
        batches = B * ((T + 7) // 8)
        print(f"B={B} L={L}: {batches} batches, {batches / 148:.2f} per CTA; cycles per batch (one warp per role, summed over CTAs / batches)")
        for r in range(5):
            tot = sum(d[8 * r:8 * r + 8])
            print(f"  {roles[r]}: total {tot / batches:7.0f}")
            for i in range(8):
                if names[r][i]:
                    print(f"     {names[r][i]:34s} {d[8 * r + i] / batches:8.0f}")


if __name__ == "__main__":
    if sys.argv[1] == "probe":
        probe()
    elif sys.argv[1] == "build":
        build_variants(sys.argv[2:])
    elif sys.argv[1] == "run":
        run_variants(sys.argv[2] if len(sys.argv) > 2 else "C2")
    else:
        timing()
