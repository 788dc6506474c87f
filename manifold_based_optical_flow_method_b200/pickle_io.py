"""The reference's third on-disk format: ``pickle.dump(obj,
bz2.BZ2File(path, 'wb'))`` for the speed maps and wave-speed arrays
(S3_compute_v_and_detection_singularity.py:136-137, S5_compute_wave_v.py:317-318, also S6:261, S7:260).

bzip2 compresses ~10 MB/s on one core: the (999, 163842) float64 speed map of one trial (1.3 GB) takes
about two minutes next to a 4.5 s solve.  ``dump`` writes the same pickle as a sequence of
independent bzip2 streams compressed by all host threads (``bz2.compress`` releases the GIL); a
multi-stream file is what ``bz2.BZ2File`` -- and therefore the reference's own
``pickle.load(bz2.BZ2File(path, 'rb'))`` (S7:218-219) -- reads back transparently.  Host-side I/O,
no GPU involved.
"""
import bz2
import os
import pickle
from concurrent.futures import ThreadPoolExecutor

CHUNK_BYTES = 8 << 20


def dump(obj, path, threads=None, compresslevel=9, chunk_bytes=CHUNK_BYTES):
    """pickle ``obj`` to ``path`` as concatenated bzip2 streams.  Returns the number of bytes written."""
    threads = threads or min(32, os.cpu_count() or 1)
    parts = []
    pickle.dump(obj, _Collector(parts, chunk_bytes), protocol=pickle.DEFAULT_PROTOCOL)
    written = 0
    with open(path, "wb") as f, ThreadPoolExecutor(max_workers=threads) as pool:
        for blob in pool.map(lambda b: bz2.compress(b, compresslevel), (bytes(p) for p in parts)):
            f.write(blob)
            written += len(blob)
    return written


def load(path):
    """What the reference does: ``pickle.load(bz2.BZ2File(path, 'rb'))``."""
    with bz2.BZ2File(path, "rb") as f:
        return pickle.load(f)


class _Collector:
    """File-like sink that cuts the pickle byte stream into chunks of about ``chunk_bytes``."""

    def __init__(self, parts, chunk_bytes):
        self.parts, self.chunk_bytes = parts, chunk_bytes

    def write(self, b):
        view = memoryview(b).cast("B")
        n = len(view)
        pos = 0
        while pos < n:
            if not self.parts or len(self.parts[-1]) >= self.chunk_bytes:
                self.parts.append(bytearray())
            take = min(n - pos, self.chunk_bytes - len(self.parts[-1]))
            self.parts[-1] += view[pos:pos + take]
            pos += take
        return n


__all__ = ["dump", "load"]
