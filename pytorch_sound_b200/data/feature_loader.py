"""Post-collate GPU feature hook for the SpeechDataset / SpeechDataLoader path.

In the reference, spectral features are either computed per item on the CPU inside DataLoader
workers (`extra_features=[(column, callable)]`, data/dataset.py:24,41-45,85-90) or by the user's
`Trainer.forward` on the collated CUDA batch (trainer.py:202).  `GpuFeatureLoader` wraps any
SpeechDataLoader-like iterable in the MAIN process: every batch (a list of tensors produced by
`pad_collate_fn`, data/dataset.py:196-228) is moved to the device the way `to_device` does
(utils/tensor.py:6-15) and the requested feature tensors are appended where `extra_features`
columns would sit — after the data columns and before the trailing mask (data/dataset.py:85-93).
"""
import inspect
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch


def _accepts_lengths(mod: Callable) -> bool:
    """Decided once from the callable's signature (nn.Module: its forward): does it take `lengths=`?"""
    fn = mod.forward if isinstance(mod, torch.nn.Module) else mod
    try:
        params = inspect.signature(fn).parameters
    except (TypeError, ValueError):
        return False
    return 'lengths' in params or any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())


class GpuFeatureLoader:
    """features: [(batch_index, module)] — module(wav (B, L)) -> (B, C, T) on the GPU.

    `mask_index` (optional): index of the wav-level mask of ones that SpeechDataset appends when
    `is_mask=True` (data/dataset.py:73-74,92-93).  When given, per-clip lengths are derived from it and
    passed to modules that accept `lengths=` so padded tails behave as per-item `extra_features` +
    zero padding would (reflect at the clip's own end, zero frames beyond it)."""

    def __init__(self, loader: Iterable, features: Sequence[Tuple[int, Callable]], device: Optional[str] = None,
                 mask_index: Optional[int] = None, prefetch: bool = False):
        """`prefetch=True`: the host->device copies and the feature kernels of batch i + 1 are enqueued on a side
        stream while the consumer still works on batch i on its own stream (the batch is handed over with an event
        wait, its tensors are marked with `record_stream`), so with a pinned-memory loader (`pin_memory=True`,
        data/dataset.py:180) the PCIe transfer — which bounds the end-to-end rate, DESIGN.md section 6 — overlaps
        the training step instead of preceding it."""
        self.loader = loader
        self.prefetch = bool(prefetch)
        self._side = None
        self.features = list(features)
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.mask_index = mask_index
        self._takes_lengths = []
        for _, mod in self.features:
            if isinstance(mod, torch.nn.Module):
                mod.to(self.device)
            self._takes_lengths.append(_accepts_lengths(mod))

    def __len__(self):
        return len(self.loader)

    def _to_device(self, batch) -> List[torch.Tensor]:
        return [x.to(self.device, non_blocking=True) if isinstance(x, torch.Tensor) else x for x in batch]

    def attach(self, batch) -> List[torch.Tensor]:
        batch = self._to_device(batch)
        lengths = None
        if self.mask_index is not None:
            lengths = batch[self.mask_index].to(torch.float32).sum(dim=-1).to(torch.int32)
        feats = []
        for (idx, mod), takes_lengths in zip(self.features, self._takes_lengths):
            wav = batch[idx]
            feats.append(mod(wav, lengths=lengths) if (lengths is not None and takes_lengths) else mod(wav))
        if self.mask_index is not None:
            m = self.mask_index % len(batch)
            return batch[:m] + feats + batch[m:]
        return batch + feats

    def __iter__(self):
        if not self.prefetch:
            for batch in self.loader:
                yield self.attach(batch)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        side = self._side

        def stage(batch):
            side.wait_stream(torch.cuda.current_stream(self.device))  # buffers the consumer freed may be reused here
            with torch.cuda.stream(side):
                out = self.attach(batch)
                ready = torch.cuda.Event()
                ready.record(side)
            return out, ready

        pending = None
        for batch in self.loader:
            nxt = stage(batch)
            if pending is not None:
                yield self._hand_over(*pending)
            pending = nxt
        if pending is not None:
            yield self._hand_over(*pending)

    def _hand_over(self, batch, ready):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        for x in batch:
            if isinstance(x, torch.Tensor) and x.is_cuda:
                x.record_stream(cur)  # allocated on the side stream, used (and later freed) on the consumer's
        return batch
