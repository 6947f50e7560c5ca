import sys, torch
sys.path.insert(0, '/root/repo')
from pytorch_sound_b200.models.transforms import LogMelSpectrogram
m = LogMelSpectrogram(22050, 80, 1024, 1024, 256, -50, 30, 0., 8000.).cuda()
xs = [torch.randn(256, 22050, device="cuda") * 0.1 for _ in range(8)]
lens = torch.randint(12000, 22051, (256,), dtype=torch.int32, device="cuda")
for name, fn in (("plain", lambda x: m(x)), ("lengths", lambda x: m(x, lengths=lens)), ("lengths+mask", lambda x: m(x, lengths=lens, frame_mask=True))):
    for x in xs: fn(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            outs = [fn(x) for x in xs]
    torch.cuda.synchronize()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(name, "%.2f us" % (e0.elapsed_time(e1) * 1e3 / 400))
