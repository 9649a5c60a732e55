"""On-disk feature formats -> device windows.

The reference keeps pre-extracted CLIP features in
  * LMDB environments whose values are `np.savez_compressed` blobs: key `features` ([T, 768] fp32) for videos
    (/root/reference/revisionllm/data/convert_h5_to_lmdb.py: dumps_npz + put), keys `token_features` / `cls_features`
    for queries (eval_nlq_negative.py:203-209), read with `np.load(io.BytesIO(dump))` (:193-197);
  * or one bare `<video>.npy` per video (:198-200, VidChapters).
and builds the overlapping windows on the host with `features[np.linspace(...)]` (:224-235).  Here the blob is decoded on
the host (zip/deflate is host work), the whole [T, 768] fp32 array crosses PCIe ONCE from a pinned staging buffer on a side
stream, and windows are gathered + cast to bf16 on the GPU (`rvl_gather_windows`).
"""
from __future__ import annotations

import io
import os
from typing import Dict, Mapping, Optional, Tuple, Union

import numpy as np
import torch

from . import scoring
from ._cabi import RvlError


def dumps_npz(arrays: Mapping[str, np.ndarray], compress: bool = True) -> bytes:
    """The writer side of the blob format (what convert_h5_to_lmdb.py stores as an LMDB value)."""
    with io.BytesIO() as w:
        (np.savez_compressed if compress else np.savez)(w, **arrays)
        return w.getvalue()


def loads_npz(blob: bytes) -> Dict[str, np.ndarray]:
    with io.BytesIO(blob) as r:
        z = np.load(r, allow_pickle=False)          # plain float arrays only: a feature store is data, never code
        return {k: z[k] for k in z.files}


class FeatureStore:
    """`kind` = 'lmdb' (needs the `lmdb` module), 'npy' (directory of <key>.npy) or 'dict' (any mapping key -> blob bytes;
    what the tests use, and a drop-in for an already opened LMDB transaction's `.get`)."""

    def __init__(self, source: Union[str, Mapping[str, bytes]], kind: Optional[str] = None):
        self.source = source
        if kind is None:
            kind = "dict" if not isinstance(source, str) else ("npy" if any(f.endswith(".npy") for f in os.listdir(source)) else "lmdb")
        self.kind = kind
        self._txn = None
        if kind == "lmdb":
            try:
                import lmdb
            except ImportError as e:                      # the module is a requirement of the reference, absent in this image
                raise RvlError("reading an LMDB feature store needs the `lmdb` module") from e
            env = lmdb.open(source, readonly=True, create=False, max_readers=4096 * 8, readahead=False)   # eval_nlq_negative.py:151
            self._txn = env.begin(buffers=True)

    def _blob(self, key: str) -> bytes:
        if self.kind == "lmdb":
            v = self._txn.get(key.encode())
        else:
            v = self.source.get(key) if hasattr(self.source, "get") else self.source[key]
        if v is None:
            raise KeyError(key)
        return bytes(v)

    def video(self, key: str) -> np.ndarray:
        """[T, 768] fp32 frame features of a movie / video."""
        if self.kind == "npy":
            return np.load(os.path.join(self.source, key + ".npy"))
        return loads_npz(self._blob(key))["features"]

    def query(self, key: str) -> Tuple[np.ndarray, np.ndarray]:
        """(token_features [Lq, 768], cls_features [768]) of a query."""
        d = loads_npz(self._blob(key))
        return d["token_features"], d["cls_features"]


class WindowLoader:
    """Pinned staging + side-stream upload + on-device window gather."""

    def __init__(self, engine, max_frames: int = 1 << 16, dim: int = 768):
        self.engine = engine
        self.dim = dim
        self._pinned = torch.empty((max_frames, dim), dtype=torch.float32).pin_memory()
        self._stream = torch.cuda.Stream(device=engine.device)
        self._copied = None                # event after the last host->device copy out of the pinned buffer

    def upload(self, features: np.ndarray) -> torch.Tensor:
        """host [T, D] (any float dtype) -> device fp32 [T, D]; the copy runs on the loader's stream and the caller's
        current stream waits for it."""
        T = features.shape[0]
        if self._copied is not None:
            self._copied.synchronize()     # the previous upload may still be reading the pinned buffer
        if T > self._pinned.shape[0]:
            self._pinned = torch.empty((T, self.dim), dtype=torch.float32).pin_memory()
        stage = self._pinned[:T]
        stage.copy_(torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32)))
        cur = torch.cuda.current_stream(self.engine.device)
        with torch.cuda.stream(self._stream):
            dev = stage.to(self.engine.device, non_blocking=True)
            self._copied = torch.cuda.Event()
            self._copied.record(self._stream)
        cur.wait_stream(self._stream)
        dev.record_stream(cur)
        return dev

    def stage1_windows(self, features: np.ndarray, clip_length: int, num_frames: int, plus_baseline: bool = False,
                       small_video: bool = False) -> torch.Tensor:
        """[W, num_frames, D] bf16 on the device: the half-overlapping windows of eval_nlq_negative.py:221-240 (optionally
        the whole-video row of --plus_baseline appended; `small_video`: the single uniformly sampled window of :213-218)."""
        T = features.shape[0]
        if small_video:
            idx = np.linspace(0, T - 1, num_frames, dtype=np.int32)[None]
        else:
            idx = scoring.stage1_windows(T, clip_length, num_frames)
            if plus_baseline:
                idx = np.concatenate([idx, np.linspace(0, T - 1, num_frames, dtype=np.int32)[None]], axis=0)
        dev = self.upload(features)
        return self.engine.gather_windows(dev, torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).to(self.engine.device))
