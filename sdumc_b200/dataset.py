"""Feature store + collate for the 4-feature SDUMC dataset.

On-disk format of the reference (toolkit/utils/read_data.py:22-49, toolkit/data/feat_data.py:171-258,
toolkit/dataloader/cmumosei.py:133-145):
    <feat_root>/<feature_name>/<utterance>.npy        float array [T, D]  (or a directory of per-frame .npy)
    label .npz with pickled dicts  train_corpus / val_corpus / test_corpus : {utterance: {'emo': ., 'val': .}}
Utterances are kept as bf16 tensors in pinned host memory; a batch is right-zero-padded per modality to the
batch maximum exactly like pad_to_maxlen_pre_modality_tensor_4 (read_data.py:223-248) — the model has no
mask, so padded frames take part in both softmaxes and the padding rule is part of the semantics.
"""
from __future__ import annotations

import os
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

STREAMS = ("audio", "text", "video", "feat4")


def read_one_feature(feature_root: str, name: str) -> np.ndarray:
    """[T, D] float array of one utterance (read_data.py:22-49)."""
    path, dpath = os.path.join(feature_root, name + ".npy"), os.path.join(feature_root, name)
    if os.path.exists(path):
        feats = [np.load(path).squeeze()]
    elif os.path.isdir(dpath):
        feats = [np.load(os.path.join(dpath, f)) for f in sorted(os.listdir(dpath))]
    else:
        raise FileNotFoundError(f"feature path or dir do not exist: {path}")
    x = np.array(feats).squeeze()
    if x.ndim == 1:
        x = x[np.newaxis, :]
    return x


def read_names_labels(label_path: str, split: str, debug: bool = False, exclude: Sequence[str] = ()):
    """(names, vals) of a split ('train' | 'val' | 'test') of the label .npz (cmumosei.py:133-145)."""
    corpus = np.load(label_path, allow_pickle=True)[f"{split}_corpus"].tolist()
    ex = set(exclude)
    names = [n for n in corpus if n not in ex]
    if debug:
        names = names[:100]
    return names, [float(corpus[n]["val"]) for n in names]


def batch_chunks(n: int, batch_size: int, rank: int = 0, world: int = 1, lockstep: bool = False) -> List[List[int]]:
    """Index lists of the batches `rank` processes, for a split of n utterances.

    Single process / scoring: the reference's batches [i*bs, (i+1)*bs) in order (its train loader does not shuffle,
    cmumosei.py:104-110), dealt round-robin to the ranks at whole-batch granularity (a sample's output depends on
    its batch's padding) and never leaving a trailing batch of size 1 (the reference model crashes on B == 1).

    lockstep=True (data-parallel training): every optimisation step is ONE global batch of world*bs consecutive
    utterances cut into `world` equal shards, so all ranks take the same number of steps with the SAME shard size
    (the all-gather / all-reduce buffers and the global normalisation B*world assume it).  The remainder after the
    last full global batch becomes one more step of floor(remainder / world) utterances per rank when that is >= 2;
    fewer than `world` (or, in that last case, 2*world) trailing utterances are dropped."""
    if world > 1 and lockstep:
        out: List[List[int]] = []
        g = world * batch_size
        full = n // g
        for s in range(full):
            lo = s * g + rank * batch_size
            out.append(list(range(lo, lo + batch_size)))
        rest = n - full * g
        per = rest // world
        if per >= 2:
            lo = full * g + rank * per
            out.append(list(range(lo, lo + per)))
        return out
    chunks = [list(range(b, min(n, b + batch_size))) for b in range(0, n, batch_size)]
    if len(chunks) > 1 and len(chunks[-1]) == 1:
        chunks[-2] += chunks[-1]
        chunks.pop()
    return [c for i, c in enumerate(chunks) if i % world == rank]


def kfold_indices(n: int, n_splits: int, seed: int = 100):
    """[(train_idx, val_idx)] * n_splits: sklearn.model_selection.KFold(n_splits, shuffle=True, random_state=seed)
    semantics (a seeded permutation cut into n_splits contiguous folds, the first n % n_splits one larger), index
    lists sorted so that batches keep the dataset order."""
    rng = np.random.RandomState(seed)
    perm = rng.permutation(n)
    sizes = np.full(n_splits, n // n_splits, dtype=int)
    sizes[: n % n_splits] += 1
    out, cur = [], 0
    for k in range(n_splits):
        val = np.sort(perm[cur:cur + sizes[k]])
        mask = np.ones(n, dtype=bool)
        mask[val] = False
        out.append((np.nonzero(mask)[0].tolist(), val.tolist()))
        cur += sizes[k]
    return out


class Store4F:
    """Four per-utterance feature lists resident in pinned host memory (bf16) + labels."""

    def __init__(self, feats: Dict[str, List[torch.Tensor]], vals: Sequence[float], names: Sequence[str]):
        self.feats, self.vals, self.names = feats, torch.tensor(list(vals), dtype=torch.float32), list(names)
        n = len(self.names)
        assert all(len(feats[s]) == n for s in STREAMS) and len(self.vals) == n
        self.dims = tuple(int(feats[s][0].shape[1]) for s in STREAMS)
        self.max_frames = tuple(max(int(x.shape[0]) for x in feats[s]) for s in STREAMS)

    def __len__(self):
        return len(self.names)

    @classmethod
    def from_disk(cls, feat_root: str, feature_names: Sequence[str], names: Sequence[str], vals: Sequence[float]):
        feats = {}
        for s, fname in zip(STREAMS, feature_names):
            root = os.path.join(feat_root, fname)
            feats[s] = [torch.from_numpy(np.ascontiguousarray(read_one_feature(root, n), dtype=np.float32)).bfloat16()
                        for n in names]
        return cls(feats, vals, names)

    @classmethod
    def synthetic(cls, n: int, dims=(1024, 4096, 1024, 4096), frames=(384, 64, 256, 64), seed: int = 1234,
                  ragged: bool = False, chunk: int = 256):
        """S0-like utterances (sdumc_b200.data.synth_batch); ragged=True draws lengths in [L/4, L]."""
        from .data import synth_batch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        feats = {s: [] for s in STREAMS}
        vals: List[float] = []
        g = torch.Generator().manual_seed(seed)
        for start in range(0, n, chunk):
            b = min(chunk, n - start)
            batch = synth_batch(b, dims, frames, seed=seed + start, device=dev)
            for s in STREAMS:
                x = batch[s].cpu()
                for i in range(b):
                    L = int(torch.randint(max(1, frames[STREAMS.index(s)] // 4), frames[STREAMS.index(s)] + 1, (1,),
                                          generator=g)) if ragged else x.shape[1]
                    feats[s].append(x[i, :L].clone())
            vals += batch["vals"].cpu().tolist()
        return cls(feats, vals, [f"synthetic_{i:06d}" for i in range(n)])

    def collate(self, idx: Sequence[int]) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, List[str]]:
        """Right-zero-pad each modality to the batch maximum and stack (read_data.py:223-248)."""
        out = {}
        for s in STREAMS:
            xs = [self.feats[s][i] for i in idx]
            L = max(int(x.shape[0]) for x in xs)
            buf = torch.zeros(len(xs), L, xs[0].shape[1], dtype=torch.bfloat16).pin_memory() \
                if torch.cuda.is_available() else torch.zeros(len(xs), L, xs[0].shape[1], dtype=torch.bfloat16)
            for j, x in enumerate(xs):
                buf[j, : x.shape[0]] = x
            out[s] = buf
        return out, self.vals[list(idx)], [self.names[i] for i in idx]

    def subset(self, idx: Sequence[int]) -> "Store4F":
        """A store over the utterances idx (shares the feature tensors): one side of a cross-validation fold."""
        idx = list(idx)
        return Store4F({s: [self.feats[s][i] for i in idx] for s in STREAMS}, self.vals[idx].tolist(),
                       [self.names[i] for i in idx])

    def batches(self, batch_size: int, rank: int = 0, world: int = 1, lockstep: bool = False) -> Iterator:
        """Collated batches of this rank (composition: batch_chunks)."""
        for c in batch_chunks(len(self), batch_size, rank, world, lockstep):
            yield self.collate(c)


class DeviceStore4F:
    """The same store resident in HBM: per modality ONE packed bf16 tensor [sum_T, D] + row
    offsets; a batch is built on the device by the collate kernel (sdumc_collate_pad) straight into the
    trainer's static input buffers - no host->device traffic per step.  CMU-MOSEI's train split is ~38 GB in
    bf16, far inside the 180 GB of a B200."""

    def __init__(self, store: Store4F, device):
        self._ids: Optional[List[int]] = None          # subset(): local position -> utterance of the packed store
        self.names, self.vals_host = store.names, store.vals
        self.dims, self.max_frames = store.dims, store.max_frames
        self.device = torch.device(device)
        self.lengths = {s: [int(x.shape[0]) for x in store.feats[s]] for s in STREAMS}     # host copy: batch maxima
        self.packed, self.offsets = {}, {}
        for s in STREAMS:
            self.packed[s] = torch.cat(store.feats[s], dim=0).to(self.device, torch.bfloat16).contiguous()
            off = torch.zeros(len(store) + 1, dtype=torch.int64)
            off[1:] = torch.cumsum(torch.tensor(self.lengths[s], dtype=torch.int64), 0)
            self.offsets[s] = off.to(self.device)
        self.vals = store.vals.to(self.device)

    @classmethod
    def from_device_batch(cls, batch: Dict[str, torch.Tensor], vals: torch.Tensor, names: Optional[List[str]] = None):
        """Store over fixed-length utterances that already sit on the device as [N, L, D] tensors (synthetic data
        generated in HBM): the packed layout is a view of the batch, no host round trip."""
        self = cls.__new__(cls)
        n = int(vals.shape[0])
        self._ids = None
        self.names = names or [f"synthetic_{i:06d}" for i in range(n)]
        self.vals_host = vals.detach().float().cpu()
        self.device = vals.device
        self.dims = tuple(int(batch[s].shape[2]) for s in STREAMS)
        self.max_frames = tuple(int(batch[s].shape[1]) for s in STREAMS)
        self.lengths = {s: [int(batch[s].shape[1])] * n for s in STREAMS}
        self.packed = {s: batch[s].to(torch.bfloat16).reshape(n * batch[s].shape[1], batch[s].shape[2]).contiguous()
                       for s in STREAMS}
        self.offsets = {s: (torch.arange(n + 1, dtype=torch.int64, device=self.device) * int(batch[s].shape[1]))
                        for s in STREAMS}
        self.vals = vals.detach().float().to(self.device)
        return self

    def __len__(self):
        return len(self.names) if self._ids is None else len(self._ids)

    def subset(self, idx: Sequence[int]) -> "DeviceStore4F":
        """A view over the utterances idx sharing the packed HBM tensors (cross-validation folds cost no memory)."""
        import copy
        sub = copy.copy(self)
        base = self._ids if self._ids is not None else list(range(len(self.names)))
        sub._ids = [base[i] for i in idx]
        ml = tuple(max(self.lengths[s][u] for u in sub._ids) for s in STREAMS)
        sub.max_frames = ml
        return sub

    def _global(self, local: Sequence[int]) -> List[int]:
        return list(local) if self._ids is None else [self._ids[i] for i in local]

    def batch_frames(self, idx: Sequence[int]) -> Tuple[int, int, int, int]:
        """Per-modality batch maximum = the padded length the reference collater would produce (idx: utterance ids
        of the packed store, as yielded by batches())."""
        if getattr(self, "_len_np", None) is None or any(len(self._len_np[s]) != len(self.lengths[s]) for s in STREAMS):
            self._len_np = {s: np.asarray(self.lengths[s], dtype=np.int64) for s in STREAMS}    # vectorised maxima
        ii = np.asarray(idx, dtype=np.int64)
        return tuple(int(self._len_np[s][ii].max()) for s in STREAMS)

    def batches(self, batch_size: int, rank: int = 0, world: int = 1, lockstep: bool = False) -> Iterator:
        """Same batch composition as Store4F.batches, yielding utterance-id lists (+ host labels, names)."""
        for c in batch_chunks(len(self), batch_size, rank, world, lockstep):
            ids = self._global(c)
            yield ids, self.vals_host[ids], [self.names[j] for j in ids]
