"""Frame decode feeding K1: N decoder processes -> one reusable pinned ring -> FramePipeline.

The reference decodes on one Python thread, one ``cap.read()`` per frame inside the DataLoader loop (reference
frameID/data.py:211-213, segment_video.py:28-45): end to end it is decode-bound at ~10^2 frames/s whatever the classifier
costs.  Hardware decode (NVDEC) is not available in this image (no nvcuvid header or library, no PyAV/ffmpeg binding other
than OpenCV's), so this is the CPU half of that row: the video's frame range is cut into W contiguous TIME RANGES, one
decoder process each (``cv2.VideoCapture`` + one seek), and every process writes the frames it decodes -- only the source
rows the resize reads, e.g. 144 of 720 -- straight into its slots of ONE shared-memory ring that the parent has registered
as pinned memory once (no per-batch ``pin_memory()``).  The parent hands full slots to ``FramePipeline.push_host`` (an
asynchronous H2D copy + the kernels) and gives a slot back to its worker when the copy's event has completed.  Each range has
its own streaming run-length encoder; the ranges' run tables are joined on the device at the end with the same pack/stitch
kernels the multi-GPU path uses (cutdet.shard.stitch_local), so the result is the table of the whole video.

Frame exactness.  A worker that starts at frame ``lo`` relies on OpenCV's frame-accurate seek.  That is checked, not assumed:
every worker but the last decodes ONE frame past its range and reports a CRC of it, which must equal the CRC of the next
worker's first frame; on a mismatch (a container whose seek is not frame accurate) ``DecodePool`` raises ``SeekMismatch`` and
the caller falls back to one sequential decoder, which reads exactly what the reference reads.

Workers import only numpy and cv2 (this module has no torch import at module level).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import queue
import zlib
from multiprocessing import shared_memory

import numpy as np


# Scheduling of the decoder processes (module defaults; DecodePool's arguments override them).  The parent needs the CPU in
# bursts -- creating the CUDA context while the workers already decode, then one short push per chunk -- and ffmpeg starts one
# decoding thread per core in EVERY worker, so W workers oversubscribe the box W times and the parent's bursts stretch (measured:
# context + weights 0.9 s alone, 2.6 s next to 8 such workers).  The workers therefore run at a lower priority, and each gets
# its share of the cores as ffmpeg threads.
WORKER_NICE = 10
DECODER_THREADS = "auto"          # "auto": max(1, cores // workers); an int; or None for ffmpeg's default (one per core)


class SeekMismatch(RuntimeError):
    """The first frame a worker decoded after its seek is not the frame that follows the previous worker's range."""


def probe_video(path: str):
    """(n_frames from the container, height, width) -- the size from a decoded frame, which is authoritative."""
    import cv2
    cap = cv2.VideoCapture(path)
    n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    ret, frame = cap.read()
    cap.release()
    if not ret:
        return 0, 0, 0
    return n, int(frame.shape[0]), int(frame.shape[1])


def _crc(a: np.ndarray) -> int:
    return zlib.crc32(memoryview(np.ascontiguousarray(a))) & 0xFFFFFFFF


def _slot_view(buf, slot: int, slot_stride: int, slot_shape: tuple) -> np.ndarray:
    return np.ndarray(slot_shape, dtype=np.uint8, buffer=buf, offset=slot * slot_stride)


def _worker_main(worker: int, path: str, shm_name: str, slot_shape: tuple, slot_stride: int, my_slots: list, lo: int, hi: int,
                 to_eof: bool, rows: np.ndarray, check_next: bool, free_q, ready_q, nice: int = 0, threads=None):
    """Decode frames [lo, hi) (or to the end of the file when ``to_eof``) into this worker's ring slots.
    Messages to the parent: ("chunk", worker, slot, n_frames, first_frame, crc_of_first_frame_or_None),
    ("done", worker, frames_decoded, crc_of_frame_hi_or_None), ("error", worker, text)."""
    shm = None
    try:
        if nice:
            try:
                os.nice(int(nice))
            except OSError:
                pass
        if threads:                             # read by OpenCV's ffmpeg backend when the capture is opened
            os.environ["OPENCV_FFMPEG_CAPTURE_OPTIONS"] = f"threads;{int(threads)}"
        import cv2
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            raise RuntimeError(f"cannot open {path}")
        if lo > 0:
            cap.set(cv2.CAP_PROP_POS_FRAMES, lo)
        shm = shared_memory.SharedMemory(name=shm_name)
        ring = {s: _slot_view(shm.buf, s, slot_stride, slot_shape) for s in my_slots}
        chunk = slot_shape[0]
        done, eof, first = 0, False, True
        while not eof and (to_eof or lo + done < hi):
            slot = free_q.get()
            if slot is None:                    # the parent is shutting down
                ring = view = None
                return
            view, k, crc = ring[slot], 0, None
            while k < chunk and (to_eof or lo + done < hi):
                ret, frame = cap.read()
                if not ret:
                    eof = True
                    break
                view[k] = frame[rows]           # only the source rows the resize reads
                if first:
                    crc, first = _crc(view[k]), False
                k += 1
                done += 1
            if k:
                ready_q.put(("chunk", worker, slot, k, lo + done - k, crc))
            else:
                free_q.put(slot)
        tail = None
        if check_next and not eof:              # one frame past my range: must be the next worker's first frame
            ret, frame = cap.read()
            if ret:
                tail = _crc(frame[rows])
        cap.release()
        ring = view = None                      # no exported buffers left when the mapping is closed
        ready_q.put(("done", worker, done, tail))
    except Exception as e:      # pragma: no cover  (reported to the parent, which raises)
        ready_q.put(("error", worker, repr(e)))
    finally:
        if shm is not None:
            shm.close()


class DecodePool:
    """Iterate over decoded chunks: ``for worker, frames, first_frame, slot in pool`` where ``frames`` is a pinned uint8 torch
    tensor view [n, len(rows), w, 3] of one ring slot (row-compacted BGR frames in decode order within the worker's range).  Call
    ``pool.release(slot, event)`` once the frames have been queued for upload: the slot goes back to its worker when the CUDA
    event has completed.  ``pool.ranges`` lists each worker's [lo, hi); ``pool.frames_decoded`` the true counts afterwards."""

    def __init__(self, path: str, rows: np.ndarray, src_h: int, src_w: int, chunk: int, n_workers: int, n_frames: int,
                 to_eof: bool = True, slots_per_worker: int = 2, pin: bool = True, start_method: str | None = None,
                 worker_nice: int | None = None, decoder_threads="default"):
        import torch
        self.path, self.chunk = path, int(chunk)
        self.rows = np.ascontiguousarray(rows, dtype=np.int64)
        n_workers = max(1, min(int(n_workers), max(1, -(-n_frames // self.chunk)))) if n_frames > 0 else 1
        self.n_workers = n_workers
        per = -(-n_frames // n_workers) if n_frames > 0 else 0
        per = -(-per // self.chunk) * self.chunk                     # whole chunks per range (the last range takes the rest)
        self.ranges = []
        for w in range(n_workers):
            lo = min(n_frames, w * per) if n_frames > 0 else 0
            hi = n_frames if w == n_workers - 1 else min(n_frames, lo + per)
            self.ranges.append((lo, hi))
        self.to_eof = bool(to_eof)
        n_slots = n_workers * slots_per_worker
        self.slot_shape = (self.chunk, len(self.rows), src_w, 3)
        self.slot_bytes = int(np.prod(self.slot_shape))
        self.slot_stride = -(-self.slot_bytes // 4096) * 4096        # every slot starts on a page: slots are pinned one by one
        self._shm = shared_memory.SharedMemory(create=True, size=n_slots * self.slot_stride)
        self._slots_np = [_slot_view(self._shm.buf, s, self.slot_stride, self.slot_shape) for s in range(n_slots)]
        self.slots = [torch.from_numpy(a) for a in self._slots_np]
        self._registered = set()
        self._pin = bool(pin)
        # fork: the workers start at once and never touch CUDA or torch (they run _worker_main: numpy + cv2 only), which is what
        # torch's own DataLoader workers rely on; "spawn" re-imports the parent's __main__ (and torch with it) in every worker
        if start_method is None:
            start_method = "fork" if "fork" in mp.get_all_start_methods() else "spawn"
        ctx = mp.get_context(start_method)
        nice = WORKER_NICE if worker_nice is None else int(worker_nice)
        threads = DECODER_THREADS if decoder_threads == "default" else decoder_threads
        if threads == "auto":
            threads = max(1, (os.cpu_count() or 1) // n_workers) if n_workers > 1 else None
        self._ready = ctx.Queue()
        self._free = [ctx.Queue() for _ in range(n_workers)]
        self._procs = []
        self._pending = []                 # (event, slot)
        self._slot_owner = {}
        for w in range(n_workers):
            slots = list(range(w * slots_per_worker, (w + 1) * slots_per_worker))
            for s in slots:
                self._slot_owner[s] = w
                self._free[w].put(s)
            lo, hi = self.ranges[w]
            last = w == n_workers - 1
            p = ctx.Process(target=_worker_main, daemon=True,
                            args=(w, path, self._shm.name, self.slot_shape, self.slot_stride, slots, lo, hi,
                                  self.to_eof and last, self.rows, not last, self._free[w], self._ready, nice, threads))
            p.start()
            self._procs.append(p)
        self.frames_decoded = [0] * n_workers
        self._first_crc = [None] * n_workers
        self._tail_crc = [None] * n_workers
        self._closed = False

    def pin(self) -> None:
        """Pin the ring from now on.  A slot is registered as pinned memory the first time a worker hands it over (its pages
        exist by then, and the other workers keep decoding meanwhile) instead of the whole ring up front, which costs ~0.3 s for
        a 400 MB ring before the first frame is scored.  ``DecodePool(..., pin=False)`` + ``pin()`` later lets the workers be
        forked BEFORE the process creates its CUDA context: forking a process that holds one is several times slower."""
        self._pin = True

    def _register(self, slot: int) -> None:
        import torch
        if self._pin and slot not in self._registered and torch.cuda.is_available():
            err = torch.cuda.cudart().cudaHostRegister(self.slots[slot].data_ptr(), self.slot_bytes, 0)
            if int(err) == 0:
                self._registered.add(slot)

    # ------------------------------------------------------------------ slots
    def release(self, slot: int, event=None) -> None:
        """The slot's frames have been handed to the GPU; ``event`` completes when the H2D copy has read them."""
        if event is None:
            self._free[self._slot_owner[slot]].put(slot)
        else:
            self._pending.append((event, slot))

    def _reap(self, block: bool = False) -> None:
        keep = []
        for i, (ev, slot) in enumerate(self._pending):
            if block and i == 0:
                ev.synchronize()
            if ev.query():
                self._free[self._slot_owner[slot]].put(slot)
            else:
                keep.append((ev, slot))
        self._pending = keep

    # ------------------------------------------------------------------ iteration
    def __iter__(self):
        running = self.n_workers
        while running:
            self._reap()
            try:
                msg = self._ready.get(timeout=0.002 if self._pending else 1.0)
            except queue.Empty:
                if self._pending:
                    self._reap(block=True)
                elif not any(p.is_alive() for p in self._procs) and self._ready.empty():
                    raise RuntimeError("decode workers exited without reporting")
                continue
            kind = msg[0]
            if kind == "error":
                self.close()
                raise RuntimeError(f"decode worker {msg[1]} failed: {msg[2]}")
            if kind == "done":
                _, w, n, tail = msg
                self.frames_decoded[w] = n
                self._tail_crc[w] = tail
                running -= 1
                continue
            _, w, slot, k, first_frame, crc = msg
            if crc is not None:
                self._first_crc[w] = crc
            self._register(slot)
            yield w, self.slots[slot][:k], first_frame, slot
        self._check_seams()

    def _check_seams(self) -> None:
        for w in range(self.n_workers - 1):
            lo, hi = self.ranges[w]
            if self.frames_decoded[w] != hi - lo:
                raise SeekMismatch(f"worker {w} decoded {self.frames_decoded[w]} frames of its range [{lo}, {hi}): the "
                                   "container's frame count is not reliable")
            nxt = self._first_crc[w + 1]
            if self._tail_crc[w] is not None and nxt is not None and self._tail_crc[w] != nxt:
                raise SeekMismatch(f"frame {hi} decoded after a seek differs from the same frame decoded in sequence")

    # ------------------------------------------------------------------ teardown
    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        for q in self._free:
            try:
                q.put(None)
            except Exception:
                pass
        for p in self._procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
        import torch
        if self._registered:
            torch.cuda.synchronize()
            for slot in sorted(self._registered):
                torch.cuda.cudart().cudaHostUnregister(self.slots[slot].data_ptr())
            self._registered.clear()
        self.slots = None
        self._slots_np = None
        try:
            self._shm.close()
            self._shm.unlink()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_workers() -> int:
    return max(1, min(8, (os.cpu_count() or 2) // 2))
