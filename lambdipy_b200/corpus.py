"""Synthetic ELF corpus (BASELINE.json configs 4 and 5).

Each file is a valid ELF64 ET_DYN modelled on what GNU ld writes (R / RX / R / RW PT_LOADs,
.dynsym/.dynstr/.rela.dyn ..., .text, .rodata, .eh_frame, .data/.bss, then the non-alloc tail:
.comment, optional .gnu.build.attributes, .debug_*, .symtab, .strtab, .shstrtab, section headers).
GNU strip accepts every file (tests/test_corpus.py checks that against the real binary), so parity
can be sampled on the synthetic corpus too.

Only the small structural pieces (headers, notes, string tables, zeroed symbol tables) are built on
the host; payload bytes of .text/.rodata/.debug_* ... are a pure function of the ARENA offset:

    byte(o) = byte (o & 7) of splitmix64(seed + (o >> 3))

generated on the device by lb2_corpus_fill (csrc/corpus.cu) or here with numpy (`payload_bytes`),
so a 100 GB corpus never exists on the host and any file can be re-materialised for checking.
"""
import struct

import numpy as np

ALIGN = 256
MASK64 = (1 << 64) - 1


# ---------------------------------------------------------------- counter-based payload (== corpus.cu)
def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def payload_bytes(seed, arena_off, length):
    """Bytes [arena_off, arena_off+length) of the payload stream."""
    if length == 0:
        return b""
    w0 = arena_off >> 3
    w1 = (arena_off + length + 7) >> 3
    with np.errstate(over="ignore"):
        words = splitmix64(np.uint64(seed & MASK64) + np.arange(w0, w1, dtype=np.uint64))
    raw = words.astype("<u8").tobytes()
    s = arena_off - (w0 << 3)
    return raw[s:s + length]


# ---------------------------------------------------------------- ELF skeleton
def _shdr(name, typ, flags, addr, off, size, link=0, info=0, align=1, entsize=0):
    return dict(name=name, type=typ, flags=flags, addr=addr, off=off, size=size, link=link, info=info, align=align, entsize=entsize)


SHT_PROGBITS, SHT_SYMTAB, SHT_STRTAB, SHT_RELA, SHT_HASH, SHT_DYNAMIC, SHT_NOTE, SHT_NOBITS, SHT_DYNSYM = 1, 2, 3, 4, 5, 6, 7, 8, 11
SHT_INIT_ARRAY, SHT_FINI_ARRAY, SHT_GNU_HASH, SHT_VERNEED, SHT_VERSYM = 14, 15, 0x6ffffff6, 0x6ffffffe, 0x6fffffff
A, W, X, M, S, I = 2, 1, 4, 0x10, 0x20, 0x40
PT_LOAD, PT_DYNAMIC, PT_NOTE, PT_EH, PT_STACK, PT_RELRO = 1, 2, 4, 0x6474e550, 0x6474e551, 0x6474e552

BUILD_ID_NOTE = struct.pack("<III", 4, 20, 3) + b"GNU\0"
GA_VERSION = b"GA$\x013a1\x00"


def _build_attr_notes(rng, text_addr, text_size):
    """Simple annobin-style version notes over pieces of .text (R9 'simple case')."""
    n = int(rng.integers(2, 9))
    cuts = np.sort(rng.integers(0, max(text_size, 1), size=2 * n))
    out = b""
    for k in range(n):
        s, e = text_addr + int(cuts[2 * k]), text_addr + int(cuts[2 * k + 1])
        out += struct.pack("<III", 8, 16, 0x100) + GA_VERSION + struct.pack("<QQ", s, e)
    return out


class FileSpec:
    """Layout of one synthetic file: total size, host-built blobs [(offset, bytes)], payload
    regions [(offset, len)] (file-relative), and bookkeeping for expectations."""
    __slots__ = ("size", "blobs", "payload", "n_sections", "kind")

    def __init__(self):
        self.size = 0
        self.blobs = []
        self.payload = []
        self.n_sections = 0
        self.kind = ""


def make_file(rng, target_size, debug_frac):
    """Build the skeleton of one ELF of roughly `target_size` bytes of which about `debug_frac`
    is dropped by strip (debug sections + .symtab + .strtab)."""
    fs = FileSpec()
    tiny = target_size < 6144
    page = 0x1000
    separate_code = (not tiny) and target_size >= 40960 and rng.random() < 0.8
    vdelta = 0 if separate_code else (0 if tiny else 0x200000)  # RW segment: vaddr - offset

    secs = []      # section dicts in file order
    blobs = []     # (off, bytes)
    payload = []   # (off, len)
    phdrs = []

    def al(v, a):
        return (v + a - 1) // a * a

    n_ph = 2 if tiny else (9 if separate_code else 7)
    cur = 64 + 56 * n_ph
    drop_budget = max(int(target_size * debug_frac), 96)
    keep_budget = max(target_size - drop_budget - (700 if tiny else 3200), 64)

    def add(name, typ, flags, size, align=1, entsize=0, link=0, info=0, data=None, fill=False, addr_delta=0, nobits=False):
        nonlocal cur
        cur = al(cur, align)
        off = cur
        addr = off + addr_delta if (flags & A) else 0
        secs.append(_shdr(name, typ, flags, addr, off, size, link, info, align, entsize))
        if not nobits:
            if data is not None:
                assert len(data) == size
                blobs.append((off, data))
            elif fill and size:
                payload.append((off, size))
            cur += size
        return len(secs)  # 1-based section index

    if tiny:
        text = max(keep_budget * 3 // 4, 16)
        data = max(keep_budget - text, 8)
        seg0 = 0
        add(".text", SHT_PROGBITS, A | X, text, 16, fill=True)
        add(".data", SHT_PROGBITS, A | W, data, 8, fill=True)
        end_alloc = cur
        phdrs.append((PT_LOAD, 7, seg0, 0, end_alloc, end_alloc, page))
        phdrs.append((PT_STACK, 6, 0, 0, 0, 0, 16))
        symtab_link_later = True
    else:
        # ---- R segment: dynamic linking tables
        k = keep_budget
        n_dynsym = int(min(max(k // 2000, 4), 20000))
        dynstr_sz = n_dynsym * 12 + 1
        n_rela = int(min(max(k // 1500, 2), 200000))
        add(".note.gnu.build-id", SHT_NOTE, A, 36, 4, data=BUILD_ID_NOTE + bytes(rng.integers(0, 256, 20, dtype=np.uint8)))
        note_off = secs[-1]["off"]
        i_hash = add(".gnu.hash", SHT_GNU_HASH, A, al(n_dynsym * 4 + 32, 8), 8, data=None)  # zeros
        i_dynsym = add(".dynsym", SHT_DYNSYM, A, n_dynsym * 24, 8, 24, info=1)            # zeros (null symbols)
        i_dynstr = add(".dynstr", SHT_STRTAB, A, dynstr_sz, 1, fill=False,
                       data=b"\0" + bytes(rng.integers(97, 123, dynstr_sz - 2, dtype=np.uint8)) + b"\0")
        add(".gnu.version", SHT_VERSYM, A, n_dynsym * 2, 2, 2, link=i_dynsym)
        add(".gnu.version_r", SHT_VERNEED, A, 64, 8, link=i_dynstr, info=1)
        add(".rela.dyn", SHT_RELA, A, n_rela * 24, 8, 24, link=i_dynsym, fill=True)
        i_relaplt = add(".rela.plt", SHT_RELA, A | I, 24 * 8, 8, 24, link=i_dynsym, info=0, fill=True)
        secs[i_hash - 1]["link"] = i_dynsym
        secs[i_dynsym - 1]["link"] = i_dynstr
        end_r = cur
        # ---- RX segment
        if separate_code:
            cur = al(cur, page)
        rx_off = cur if separate_code else 0
        text_sz = max(int(k * 0.62), 64)
        add(".init", SHT_PROGBITS, A | X, 27, 4, fill=True)
        i_plt = add(".plt", SHT_PROGBITS, A | X, 16 * 9, 16, 16, fill=True)
        add(".text", SHT_PROGBITS, A | X, text_sz, 64 if text_sz > 65536 else 16, fill=True)
        text_addr = secs[-1]["addr"]
        add(".fini", SHT_PROGBITS, A | X, 13, 4, fill=True)
        end_rx = cur
        # ---- R segment 2
        if separate_code:
            cur = al(cur, page)
        ro_off = cur
        add(".rodata", SHT_PROGBITS, A, max(int(k * 0.2), 16), 32, fill=True)
        ehh = al(max(int(k * 0.01), 12), 4)
        add(".eh_frame_hdr", SHT_PROGBITS, A, ehh, 4, fill=True)
        ehh_off = secs[-1]["off"]
        add(".eh_frame", SHT_PROGBITS, A, al(max(int(k * 0.07), 24), 8), 8, fill=True)
        end_ro = cur
        # ---- RW segment
        if separate_code:
            cur = al(cur, page) + (page - 0x250)  # ld puts RELRO data at the end of a page
            vd = 0x1000
        else:
            cur = al(cur, 8)
            vd = vdelta
        rw_off = cur
        add(".init_array", SHT_INIT_ARRAY, A | W, 8, 8, 8, fill=True, addr_delta=vd)
        add(".fini_array", SHT_FINI_ARRAY, A | W, 8, 8, 8, fill=True, addr_delta=vd)
        i_dyn = add(".dynamic", SHT_DYNAMIC, A | W, 0x1f0, 8, 16, link=i_dynstr, addr_delta=vd)  # zeros = DT_NULL
        dyn_off = secs[-1]["off"]
        add(".got", SHT_PROGBITS, A | W, 0x28, 8, 8, addr_delta=vd)
        relro_end = cur
        i_gotplt = add(".got.plt", SHT_PROGBITS, A | W, 0x60, 8, 8, addr_delta=vd)
        add(".data", SHT_PROGBITS, A | W, max(int(k * 0.06), 16), 32, fill=True, addr_delta=vd)
        end_rw_file = cur
        bss_sz = int(rng.integers(8, 1 << 16))
        add(".bss", SHT_NOBITS, A | W, bss_sz, 32, addr_delta=vd, nobits=True)
        end_rw_mem = secs[-1]["addr"] + bss_sz
        secs[i_relaplt - 1]["info"] = i_gotplt
        if separate_code:
            phdrs.append((PT_LOAD, 4, 0, 0, end_r, end_r, page))
            phdrs.append((PT_LOAD, 5, rx_off, rx_off, end_rx - rx_off, end_rx - rx_off, page))
            phdrs.append((PT_LOAD, 4, ro_off, ro_off, end_ro - ro_off, end_ro - ro_off, page))
        else:
            phdrs.append((PT_LOAD, 5, 0, 0, end_ro, end_ro, 0x200000 if vdelta else page))
        phdrs.append((PT_LOAD, 6, rw_off, rw_off + vd, end_rw_file - rw_off, end_rw_mem - (rw_off + vd), 0x200000 if (vdelta and not separate_code) else page))
        phdrs.append((PT_DYNAMIC, 6, dyn_off, dyn_off + vd, 0x1f0, 0x1f0, 8))
        phdrs.append((PT_NOTE, 4, note_off, note_off, 36, 36, 4))
        phdrs.append((PT_EH, 4, ehh_off, ehh_off, ehh, ehh, 4))
        phdrs.append((PT_STACK, 6, 0, 0, 0, 0, 16))
        phdrs.append((PT_RELRO, 4, rw_off, rw_off + vd, relro_end - rw_off, relro_end - rw_off, 1))
        assert len(phdrs) == n_ph, (len(phdrs), n_ph)

    # ---- non-alloc tail
    if not tiny or rng.random() < 0.5:
        comment = b"GCC: (GNU) 13.3.0 lambdipy-b200 synthetic\0"
        add(".comment", SHT_PROGBITS, M | S, len(comment), 1, 1, data=comment)
    if not tiny and rng.random() < 0.5:
        notes = _build_attr_notes(rng, text_addr, text_sz)
        add(".gnu.build.attributes", SHT_NOTE, 0, len(notes), 4, data=notes)
        secs[-1]["addr"] = end_rw_mem + 0x1000  # annobin sections carry a bogus address
    sym_bytes = max(min(drop_budget // 5, 24 * 400000) // 24 * 24, 48)
    str_bytes = max(min(drop_budget // 6, 16 << 20), 16)
    dbg = max(drop_budget - sym_bytes - str_bytes, 0)
    if dbg >= 64:
        parts = [(".debug_aranges", 0.02, 16), (".debug_info", 0.42, 1), (".debug_abbrev", 0.04, 1), (".debug_line", 0.2, 1),
                 (".debug_str", 0.22, 1), (".debug_loclists", 0.07, 1), (".debug_rnglists", 0.03, 1)]
        if dbg < 512:
            parts = [(".debug_info", 0.7, 1), (".debug_str", 0.3, 1)]
        for nm, fr, algn in parts:
            sz = int(dbg * fr)
            if sz:
                add(nm, SHT_PROGBITS, (M | S) if nm == ".debug_str" else 0, sz, algn, 1 if nm == ".debug_str" else 0, fill=True)
    n_sec_before = len(secs)
    i_symtab = add(".symtab", SHT_SYMTAB, 0, sym_bytes, 8, 24, link=n_sec_before + 2, info=1)  # zeros: null symbols
    strtab = None
    add(".strtab", SHT_STRTAB, 0, str_bytes, 1, fill=False)
    st_off = secs[-1]["off"]
    blobs.append((st_off, b"\0"))
    if str_bytes > 2:
        payload.append((st_off + 1, str_bytes - 2))  # payload bytes; the table still ends in NUL (zero arena)
    names = [s["name"] for s in secs] + [".shstrtab"]
    shstr = b"\0"
    name_off = {}
    for nm in names:
        if nm not in name_off:
            name_off[nm] = len(shstr)
            shstr += nm.encode() + b"\0"
    add(".shstrtab", SHT_STRTAB, 0, len(shstr), 1, data=shstr)
    cur = al(cur, 8)
    shoff = cur
    table = bytearray(64)
    for s in secs:
        table += struct.pack("<IIQQQQIIQQ", name_off[s["name"]], s["type"], s["flags"], s["addr"], s["off"], s["size"],
                             s["link"], s["info"], s["align"], s["entsize"])
    blobs.append((shoff, bytes(table)))
    total = shoff + len(table)
    eh = struct.pack("<16sHHIQQQIHHHHHH", b"\x7fELF\x02\x01\x01" + b"\0" * 9, 3, 62, 1, 0, 64, shoff, 0, 64, 56, len(phdrs), 64,
                     len(secs) + 1, len(secs))
    ph = b"".join(struct.pack("<IIQQQQQQ", t, fl, off, va, va, fsz, msz, algn) for (t, fl, off, va, fsz, msz, algn) in phdrs)
    blobs.append((0, eh + ph))
    fs.size = total
    fs.blobs = sorted(blobs)
    fs.payload = payload
    fs.n_sections = len(secs) + 1
    fs.kind = "tiny" if tiny else ("sepcode" if separate_code else "compact")
    return fs


class Corpus:
    """A seeded corpus: file sizes log-uniform in [min_size, max_size], dropped fraction U(0.05, 0.8)."""

    def __init__(self, n_files, seed=0xB200, min_size=1 << 10, max_size=128 << 20, rank=0, world=1, max_total=None):
        self.seed = int(seed)
        rng = np.random.default_rng([self.seed, 7])
        sizes = np.exp(rng.uniform(np.log(min_size), np.log(max_size), size=n_files)).astype(np.int64)
        fracs = rng.uniform(0.05, 0.8, size=n_files)
        if max_total is not None:
            # keep the log-uniform shape, stop once the declared cap would be exceeded
            keep = np.cumsum(sizes) <= max_total
            sizes, fracs = sizes[keep], fracs[keep]
        from .sharding import shard_indices
        self.global_index = shard_indices(sizes, rank, world)  # size-sorted round-robin
        self.files = []
        for gi in self.global_index:
            frng = np.random.default_rng([self.seed, 11, int(gi)])
            self.files.append(make_file(frng, int(sizes[gi]), float(fracs[gi])))
        self.sizes = np.array([f.size for f in self.files], dtype=np.uint64)
        off = np.zeros(len(self.files) + 1, dtype=np.uint64)
        if len(self.files):
            np.cumsum((self.sizes + np.uint64(ALIGN - 1)) // np.uint64(ALIGN) * np.uint64(ALIGN), out=off[1:])
        self.off = off
        self.arena_bytes = int(off[-1]) + ALIGN

    def __len__(self):
        return len(self.files)

    @property
    def total_bytes(self):
        return int(self.sizes.sum())

    def fill_regions(self):
        """Arena-absolute payload regions [(offset, len)] as a (n,2) uint64 array."""
        regs = []
        for f, base in zip(self.files, self.off[:-1]):
            b = int(base)
            regs.extend((b + o, l) for o, l in f.payload)
        return np.array(regs, dtype=np.uint64).reshape(-1, 2)

    def blob_table(self):
        """All host-built blobs concatenated: (data bytes, dst_off[], src_off[], len[])."""
        data = bytearray()
        dst, src, ln = [], [], []
        for f, base in zip(self.files, self.off[:-1]):
            b = int(base)
            for o, blob in f.blobs:
                pad = (-len(data)) % 16
                data += b"\0" * pad
                dst.append(b + o)
                src.append(len(data))
                ln.append(len(blob))
                data += blob
        return (bytes(data), np.array(dst, dtype=np.uint64), np.array(src, dtype=np.uint64), np.array(ln, dtype=np.uint64))

    def materialize(self, i):
        """File i exactly as it sits in the device arena, built on the host."""
        f = self.files[i]
        base = int(self.off[i])
        buf = bytearray(f.size)
        for o, l in f.payload:
            buf[o:o + l] = payload_bytes(self.seed, base + o, l)
        for o, blob in f.blobs:
            buf[o:o + len(blob)] = blob
        return bytes(buf)
