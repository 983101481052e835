"""Oracle: per-frame scores -> segments -> CSV.  TEST INFRASTRUCTURE ONLY.

Restates ``Segmentation`` (reference frameID/segmentation.py:26-196) over plain
python lists / numpy scalars:

  * ``run_table``           -- __init__ (segmentation.py:35-60): per-frame max logit and
    argmax (first index wins ties, as torch.max does on CPU), run-length encoding,
    per-run mean of the max logit in float32.
  * ``glue_orphans``        -- segmentation.py:91-166 with ``_find_orphans`` (:12-17) and
    ``_update_neighbor`` (:69-89) including the mean-update expression
    ``(m_n*l_n + m_o*l_o) / l_n + l_o`` evaluated exactly as written (divide by the
    neighbour's length only, then add the orphan's length), in float32.
  * ``combine_adjacent``    -- segmentation.py:168-183.
  * ``csv_bytes``           -- segmentation.py:185-196: ``start_frame,label`` rows, CRLF.

One deliberate pin: the reference picks the least confident orphan with
``torch.argsort(...)[0]``, which is NOT a stable sort, so among orphans whose
float32 means are exactly equal its choice is unspecified (it depends on the
torch build and the CPU's vector width).  The oracle -- and the CUDA kernel --
choose the LOWEST run index among exact ties.  The golden vectors recorded from
the reference (tests/golden/segmentation_kat.npz) contain no exact ties.

Float note: the reference's per-run mean is torch's float32 ``.mean()``; this
oracle uses the same call when torch is importable (bit-identical), else a
float64 sum rounded to float32.
"""
from __future__ import annotations

import numpy as np

TYPE_NAMES = {0: "a22", 1: "ez", 2: "b"}          # frameID/data.py:116
BLANK = 2


def max_and_argmax(scores: np.ndarray):
    """torch.max(scores, dim=1): value and FIRST index of the maximum."""
    scores = np.asarray(scores, dtype=np.float32)
    return scores.max(axis=1), scores.argmax(axis=1).astype(np.int64)


def _mean_f32(x: np.ndarray) -> np.float32:
    try:
        import torch
        return np.float32(torch.from_numpy(np.ascontiguousarray(x)).mean().item())
    except ImportError:  # pragma: no cover
        return np.float32(np.sum(x, dtype=np.float64) / x.size)


def run_table(scores: np.ndarray) -> dict:
    """The five columns of ``Segmentation.te`` right after construction."""
    top, lab = max_and_argmax(scores)
    n = lab.shape[0]
    ends = np.flatnonzero(lab[1:] != lab[:-1]).astype(np.int64)
    ends = np.concatenate([ends, np.array([n - 1], dtype=np.int64)])
    starts = np.concatenate([np.zeros(1, dtype=np.int64), ends[:-1] + 1])
    means = np.array([_mean_f32(top[s:e + 1]) for s, e in zip(starts, ends)], dtype=np.float32)
    return {
        "end_frames": ends,
        "frame_types": lab[ends],
        "run_lengths": ends - starts + 1,
        "start_frames": starts,
        "score_means": means,
    }


def run_table_from_labels(labels: np.ndarray, top: np.ndarray) -> dict:
    """Same table from precomputed (label, max logit) columns."""
    lab = np.asarray(labels).astype(np.int64)
    top = np.asarray(top, dtype=np.float32)
    n = lab.shape[0]
    ends = np.flatnonzero(lab[1:] != lab[:-1]).astype(np.int64)
    ends = np.concatenate([ends, np.array([n - 1], dtype=np.int64)])
    starts = np.concatenate([np.zeros(1, dtype=np.int64), ends[:-1] + 1])
    means = np.array([_mean_f32(top[s:e + 1]) for s, e in zip(starts, ends)], dtype=np.float32)
    return {"end_frames": ends, "frame_types": lab[ends], "run_lengths": ends - starts + 1,
            "start_frames": starts, "score_means": means}


class _Runs:
    """Mutable list-of-columns view used by the two smoothing passes."""

    def __init__(self, te: dict):
        self.start = [int(v) for v in te["start_frames"]]
        self.end = [int(v) for v in te["end_frames"]]
        self.length = [int(v) for v in te["run_lengths"]]
        self.kind = [int(v) for v in te["frame_types"]]
        self.mean = [np.float32(v) for v in te["score_means"]]

    def absorb(self, orphan: int, neighbour: int) -> None:
        """_update_neighbor (segmentation.py:69-89), float32 step by step."""
        if orphan < neighbour:
            self.start[neighbour] = self.start[orphan]
        else:
            self.end[neighbour] = self.end[orphan]
        ln = np.float32(self.length[neighbour])
        lo = np.float32(self.length[orphan])
        num = np.float32(np.float32(self.mean[neighbour] * ln) + np.float32(self.mean[orphan] * lo))
        self.mean[neighbour] = np.float32(np.float32(num / ln) + lo)
        self.length[neighbour] = self.end[neighbour] - self.start[neighbour] + 1

    def drop(self, i: int) -> None:
        for col in (self.start, self.end, self.length, self.kind, self.mean):
            del col[i]

    def is_orphan(self, i: int, k_real: int, k_blank: int) -> bool:
        if self.kind[i] != BLANK:
            return self.length[i] < k_real
        return self.length[i] < k_blank

    def table(self) -> dict:
        return {
            "end_frames": np.array(self.end, dtype=np.int64),
            "frame_types": np.array(self.kind, dtype=np.int64),
            "run_lengths": np.array(self.length, dtype=np.int64),
            "start_frames": np.array(self.start, dtype=np.int64),
            "score_means": np.array(self.mean, dtype=np.float32),
        }


def glue_orphans(te: dict, real_threshold: int = 100, blank_threshold: int = 10) -> dict:
    r = _Runs(te)
    while True:
        orphans = [i for i in range(len(r.start)) if r.is_orphan(i, real_threshold, blank_threshold)]
        if not orphans:
            break
        # least confident, lowest index on ties; a NaN mean sorts after every number, where torch.argsort puts it
        target = min(orphans, key=lambda i: (bool(np.isnan(r.mean[i])), r.mean[i], i))
        last = len(r.start) - 1
        if target == 0:
            if last == 0:
                raise IndexError("index 1 is out of bounds for dimension 0 with size 1")
            r.absorb(0, 1)
        elif target == last:
            r.absorb(target, target - 1)
        elif r.length[target - 1] > r.length[target + 1]:
            r.absorb(target, target - 1)
        else:
            r.absorb(target, target + 1)
        r.drop(target)
    return r.table()


def combine_adjacent(te: dict) -> dict:
    r = _Runs(te)
    i = 0
    while i < len(r.start) - 1:
        # the reference always restarts from the first matching pair; after merging
        # pair (i, i+1) the merged run sits at index i, so scanning on from i is the same.
        if r.kind[i] == r.kind[i + 1]:
            r.absorb(i, i + 1)
            r.drop(i)
        else:
            i += 1
    return r.table()


def csv_bytes(te: dict) -> bytes:
    rows = [f"{int(s)},{TYPE_NAMES[int(t)]}\r\n" for s, t in zip(te["start_frames"], te["frame_types"])]
    return "".join(rows).encode("ascii")


def segment(scores: np.ndarray, real_threshold: int = 100, blank_threshold: int = 10):
    """scores -> (initial table, glued table, combined table, csv bytes): the CLI's sequence
    (segment_video.py:62-77)."""
    t0 = run_table(scores)
    t1 = glue_orphans(t0, real_threshold, blank_threshold)
    t2 = combine_adjacent(t1)
    return t0, t1, t2, csv_bytes(t2)


def stitch_tables(shards: list, frame_offsets: list) -> dict:
    """Join the run tables of consecutive time shards (LOCAL frame numbers, each with a float64 ``score_sums``
    column) into the table of the whole sequence: what ``Segmentation.__init__`` (segmentation.py:35-60) would
    have built from the concatenated scores.  Runs that meet at a shard edge with the same type are merged --
    sums and lengths add, so the mean is that of the joined slice.  Checker for cutdet_stitch_shards."""
    end, start, length, kind, sums = [], [], [], [], []
    for te, off in zip(shards, frame_offsets):
        for i in range(len(te["end_frames"])):
            e, s = int(te["end_frames"][i]) + off, int(te["start_frames"][i]) + off
            k, sm = int(te["frame_types"][i]), float(te["score_sums"][i])
            if i == 0 and kind and kind[-1] == k:
                end[-1] = e
                length[-1] = e - start[-1] + 1
                sums[-1] += sm
            else:
                end.append(e); start.append(s); length.append(e - s + 1); kind.append(k); sums.append(sm)
    return {
        "end_frames": np.array(end, dtype=np.int64),
        "frame_types": np.array(kind, dtype=np.int64),
        "run_lengths": np.array(length, dtype=np.int64),
        "start_frames": np.array(start, dtype=np.int64),
        "score_means": np.array([np.float32(s / l) for s, l in zip(sums, length)], dtype=np.float32),
        "score_sums": np.array(sums, dtype=np.float64),
    }
