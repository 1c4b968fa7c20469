"""HBM-resident batches: put a Corpus (or a list of host files) into a device arena and drive
lb2_strip_device_async / lb2_batch_results on it.  Used by bench.py and the GPU tests."""
import ctypes as C

import numpy as np

from . import _native as N


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


class DeviceBatch:
    def __init__(self, ctx, off, sizes, out_slack=None, chunk_bytes=None):
        """off: uint64[n+1] arena offsets (multiples of 256), sizes: uint64[n] exact sizes.
        chunk_bytes: keep only the INPUT resident and stream the output through two slots of that
        many bytes (+ slack) -- for shards whose input + output exceed HBM (lb2_strip_device_chunked)."""
        self.ctx = ctx
        self.n = len(sizes)
        self.off = np.ascontiguousarray(off, dtype=np.uint64)
        self.sizes = np.ascontiguousarray(sizes if self.n else np.zeros(1), dtype=np.uint64)
        self.in_bytes = int(self.off[-1]) + 256
        self.chunk_bytes = chunk_bytes
        if chunk_bytes is None:
            self.out_cap = self.in_bytes + self.n * 4096 + (16 << 20) if out_slack is None else self.in_bytes + out_slack
            self.d_out = None
        else:
            # a chunk holds whole files: at least the largest file, plus growth slack per file
            biggest = int(((self.sizes + np.uint64(255)) // np.uint64(256) * np.uint64(256)).max()) if self.n else 0
            self.chunk_bytes = max(int(chunk_bytes), biggest)
            self.slot_cap = self.chunk_bytes + min(self.n, 1 << 16) * 4096 + (16 << 20)
            self.out_cap = 2 * self.slot_cap
        self.d_in = ctx.dev_alloc(self.in_bytes)
        self.d_out = ctx.dev_alloc(self.out_cap)
        self.out_off = np.zeros(self.n + 1, dtype=np.uint64)
        self.out_sizes = np.zeros(max(self.n, 1), dtype=np.uint64)
        self.status = np.zeros(max(self.n, 1), dtype=np.int32)

    @classmethod
    def from_corpus(cls, ctx, corpus, chunk_bytes=None):
        b = cls(ctx, corpus.off, corpus.sizes, chunk_bytes=chunk_bytes)
        ctx.check(ctx.lib.lb2_memset_d(ctx.h, b.d_in, 0, b.in_bytes))
        regs = corpus.fill_regions()
        if len(regs):
            arr = (N.FillRegion * len(regs))()
            flat = np.frombuffer(arr, dtype=np.uint64)
            flat[:] = regs.reshape(-1)
            ctx.check(ctx.lib.lb2_corpus_fill(ctx.h, b.d_in, arr, len(regs), C.c_uint64(corpus.seed), None))
        data, dst, src, ln = corpus.blob_table()
        if len(dst):
            ctx.check(ctx.lib.lb2_corpus_scatter(ctx.h, b.d_in, data, len(data), _u64p(dst), _u64p(src), _u64p(ln), len(dst)))
        return b

    @classmethod
    def from_blobs(cls, ctx, blobs):
        sizes = np.array([len(x) for x in blobs], dtype=np.uint64)
        off = np.zeros(len(blobs) + 1, dtype=np.uint64)
        if len(blobs):
            np.cumsum((sizes + np.uint64(255)) // np.uint64(256) * np.uint64(256), out=off[1:])
        b = cls(ctx, off, sizes)
        host = bytearray(int(off[-1]) + 256)
        for o, x in zip(off[:-1], blobs):
            host[int(o):int(o) + len(x)] = x
        ctx.h2d(b.d_in, bytes(host), len(host))
        return b

    def strip_async(self, flags=0, stream=None):
        self.ctx.check(self.ctx.lib.lb2_strip_device_async(self.ctx.h, self.d_in, _u64p(self.off), _u64p(self.sizes), self.n,
                                                           self.d_out, self.out_cap, flags, stream))

    def results(self):
        st = N.Stats()
        self.ctx.check(self.ctx.lib.lb2_batch_results(self.ctx.h, _u64p(self.out_off), _u64p(self.out_sizes),
                                                      self.status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st)))
        return st.as_dict()

    def strip_chunked(self, flags=0, stream=None, on_chunk=None):
        """One pass over the shard with the output streamed through the two-slot ring.  on_chunk(chunk,
        first_file, n_files, d_slot, out_off, out_sizes, status) is called while the chunk's slot is valid."""
        st = N.Stats()
        cb = None
        if on_chunk is not None:
            def _cb(user, chunk, f0, n, d_slot, ooff, osz, stat, cst):
                on_chunk(chunk, f0, n, d_slot, np.ctypeslib.as_array(ooff, (n + 1,)), np.ctypeslib.as_array(osz, (n,)),
                         np.ctypeslib.as_array(stat, (n,)))
                return 0
            cb = N.CHUNK_FN(_cb)
        self.ctx.check(self.ctx.lib.lb2_strip_device_chunked(
            self.ctx.h, self.d_in, _u64p(self.off), _u64p(self.sizes), self.n, self.d_out, self.slot_cap, self.chunk_bytes, flags, stream,
            C.cast(cb, C.c_void_p) if cb else None, None, _u64p(self.out_sizes), self.status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st)))
        return st.as_dict()

    def read_input(self, i):
        n = int(self.sizes[i])
        buf = C.create_string_buffer(n)
        self.ctx.d2h(buf, self.d_in + int(self.off[i]), n)
        return buf.raw

    def read_output(self, i):
        n = int(self.out_sizes[i])
        buf = C.create_string_buffer(n)
        self.ctx.d2h(buf, self.d_out + int(self.out_off[i]), n)
        return buf.raw

    def read_input_arena(self, host_ptr):
        self.ctx.d2h(host_ptr, self.d_in, int(self.off[-1]))

    def close(self):
        if self.d_in:
            self.ctx.dev_free(self.d_in)
            self.ctx.dev_free(self.d_out)
            self.d_in = self.d_out = None
