"""On-disk artefacts of ``everyvoice preprocess`` written from batched device results
(SURVEY.md section 8f, N2).

The reference writes one tiny file per utterance and step from its worker processes
(``save_tensor`` / ``save_wav``, everyvoice/preprocessor/helpers.py:23-44) and later re-reads and
re-writes every energy / pitch file serially to normalise it (``normalize_stats``,
preprocessor.py:453-490, "this function is IO bound").  Here a whole batch comes back from the
device in ONE copy per array, already normalised, and a thread pool serialises the per-utterance
files.  Names, directory layout and formats are the reference's:

* ``{save_dir}/{folder}/{basename}--{speaker}--{language}--{fn}``      (``create_path``, preprocessor.py:502-508)
* ``spec/...--spec-{sampling_rate}-{spec_type}.pt``: ``[F, T]`` float32   (preprocessor.py:892-896, 917-928)
* ``energy/...--energy.pt`` / ``pitch/...--pitch.pt``: ``[T]`` or ``[P]`` float32 (:632-651, 653-670)
* ``audio/...--audio-{sampling_rate}.wav``: mono PCM16                    (:510-584, helpers.py:31-44)
* ``stats.json``: ``{"energy": {...}, "pitch": {...}}`` merged into an existing file (fs2/cli/preprocess.py:44-77)

Every ``.pt`` file is a plain ``torch.save`` of a contiguous CPU tensor, so the reference's
consumers (``torch.load(path, weights_only=True)`` in fs2/dataset.py:56-60, hfgl/dataset.py:60-101)
read it unchanged.  The one deliberate difference: the reference saves the ``[:, :T]`` VIEW of the
``[F, T + 1]`` transform output, which serialises the dropped last frame as well; these files hold
exactly ``F * T`` values.
"""

from __future__ import annotations

import json
import wave
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch


def create_path(save_dir, item: dict, folder: str, fn: str, sep: str = "--") -> Path:
    """Reference: Preprocessor.create_path, preprocessor.py:502-508."""
    return Path(save_dir) / folder / sep.join([item["basename"], item["speaker"], item["language"], fn])


def spec_filename(sampling_rate: int, spec_type) -> str:
    return f"spec-{sampling_rate}-{getattr(spec_type, 'value', spec_type)}.pt"


class ArtefactWriter:
    """Writes the per-utterance files of one or more batches.  ``close()`` (or the context manager)
    waits for the pool; exceptions of the workers are re-raised there."""

    def __init__(self, save_dir, sep: str = "--", workers: int = 8):
        self.save_dir = Path(save_dir)
        self.sep = sep
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(workers)))
        self._futures: list = []
        self._made: set = set()
        self.files_written = 0

    # -- helpers ------------------------------------------------------------------------------
    def path(self, item: dict, folder: str, fn: str) -> Path:
        return create_path(self.save_dir, item, folder, fn, self.sep)

    def _mkdir(self, folder: str):
        if folder not in self._made:
            (self.save_dir / folder).mkdir(parents=True, exist_ok=True)
            self._made.add(folder)

    def _submit(self, fn, *args):
        self._futures.append(self._pool.submit(fn, *args))
        self.files_written += 1

    @staticmethod
    def _save(t: torch.Tensor, path: Path):
        torch.save(t, path)

    @staticmethod
    def _save_wav(pcm: np.ndarray, path: Path, sr: int):
        with wave.open(str(path), "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(int(sr))
            w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())

    # -- batched writers ------------------------------------------------------------------------
    def write_specs(self, items, feats, sampling_rate: int, spec_type):
        """``feats``: a ``RaggedFeatures`` (device or host).  One D2H copy of the packed ``[sum T, F]`` array, then
        one contiguous ``[F, T_b]`` tensor per utterance, as process_spec saves it."""
        self._mkdir("spec")
        fn = spec_filename(sampling_rate, spec_type)
        host = feats.spec.detach().to("cpu")  # ONE copy for the batch
        off = np.asarray(feats.frame_offsets, dtype=np.int64)
        if len(items) != len(off) - 1:
            raise ValueError("items and features disagree on the number of utterances")
        for b, item in enumerate(items):
            rows = host[int(off[b]) : int(off[b + 1])]
            if feats.is_complex:
                rows = torch.view_as_complex(rows.view(rows.shape[0], feats.n_rows, 2))
            # an own, dense [F, T] tensor: torch.save serialises the WHOLE storage a view points into
            out = torch.empty((rows.shape[1], rows.shape[0]), dtype=rows.dtype)
            out.copy_(rows.transpose(0, 1))
            self._submit(self._save, out, self.path(item, "spec", fn))

    def write_ragged(self, items, values: torch.Tensor, offsets, folder: str, fn: str):
        """Packed per-utterance vectors (frame- or phone-level energy / pitch) -> ``{folder}/...--{fn}``."""
        self._mkdir(folder)
        host = values.detach().to("cpu")
        off = np.asarray(offsets.cpu() if torch.is_tensor(offsets) else offsets, dtype=np.int64)
        if len(items) != len(off) - 1:
            raise ValueError("items and offsets disagree on the number of utterances")
        for b, item in enumerate(items):
            self._submit(self._save, host[int(off[b]) : int(off[b + 1])].clone(), self.path(item, folder, fn))

    def write_energy(self, items, values, offsets):
        self.write_ragged(items, values, offsets, "energy", "energy.pt")

    def write_pitch(self, items, values, offsets):
        self.write_ragged(items, values, offsets, "pitch", "pitch.pt")

    def write_audio(self, items, processed):
        """``processed``: the ``ProcessedAudio`` of ``process_audio_batch(out_dtype=torch.int16)``; items are
        indexed like its INPUT list (skipped utterances get no file, like the reference)."""
        if processed.samples.dtype != torch.int16:
            raise ValueError("write_audio expects PCM16 samples (process_audio_batch(..., out_dtype=torch.int16))")
        self._mkdir("audio")
        host = processed.samples.detach().to("cpu").numpy()
        off = processed.offsets
        fn = f"audio-{processed.sr}.wav"
        for j, i in enumerate(processed.kept):
            self._submit(self._save_wav, host[int(off[j]) : int(off[j + 1])], self.path(items[i], "audio", fn),
                         processed.sr)

    def write_stats(self, stats: dict):
        """``stats.json`` as fs2/cli/preprocess.py:68-77 writes it: merged over an existing file."""
        self.save_dir.mkdir(parents=True, exist_ok=True)
        path = self.save_dir / "stats.json"
        previous = {}
        if path.exists():
            with open(path, "r", encoding="utf8") as f:
                previous = json.load(f)
        with open(path, "w", encoding="utf8") as f:
            json.dump({**previous, **stats}, f)
        return path

    # -- lifetime -------------------------------------------------------------------------------
    def flush(self):
        futures, self._futures = self._futures, []
        for fu in futures:
            fu.result()

    def close(self):
        try:
            self.flush()
        finally:
            self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def load_ragged(paths, device=None, workers: int = 8):
    """The reading side (``torch.load(path, weights_only=True)`` per file, as process_energy / compute_stats /
    the datasets do) for a list of 1-D tensors: returns ``(packed tensor, offsets)``; the pool overlaps the
    small reads, one H2D copy moves the batch."""
    with ThreadPoolExecutor(max_workers=max(1, int(workers))) as pool:
        tensors = list(pool.map(lambda p: torch.load(p, weights_only=True).reshape(-1), paths))
    lens = np.array([t.numel() for t in tensors], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    packed = torch.cat(tensors) if tensors else torch.zeros(0)
    return (packed.to(device, non_blocking=True) if device is not None else packed), off
