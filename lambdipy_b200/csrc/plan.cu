// plan.cu -- the ELF "plan" kernel: one 128-thread CTA per file parses Ehdr / Shdr / Phdr / .shstrtab out
// of the HBM input arena with coalesced 16-byte loads, decides keep/drop per section, lays the stripped
// file out exactly as GNU strip (Binutils 2.42) would, regenerates .shstrtab, the section-header table,
// the program-header table and merged build-attribute notes into a per-file scratch slot, and emits the
// list of (src,dst,len) tiles the compaction kernel executes.
//
// The kernel is latency-bound (a few KB of headers per file, long dependent chains), so the work is
// organised to shorten the critical path rather than to move bytes:
//   * one thread per section header / program header for everything that is independent per entry
//     (input gate, keep/drop, name hashes, segment membership, new headers);
//   * after the section order is known, the three long independent jobs run on different warps at
//     the same time:  warp 0 program headers + LOAD layout (R10-R12), warp 1 build-attribute note
//     merging (R9), warp 2 the tail-merged .shstrtab (R6);
//   * the sequential parts that remain (LOAD cursor, non-alloc packing, note range inheritance) touch a
//     handful of values each; the note sorts -- glibc's merge sort restated level by level -- compare
//     register-resident keys and prefetch the next element of both runs;
//   * one launch for every file: the note workspace (8 KB of note bytes, 320 notes) aliases the extent arrays.
//
// Replaces, per file, what the reference delegates to the external `strip` binary:
// /root/reference/lambdipy/project_build.py:260.  Rules R1..R12: /root/repo/SURVEY.md 8(c).
// This is product code: it shares nothing with oracle/ (an independent CPU restatement used only
// by the tests to check this kernel).
#include "lb2_common.cuh"
#ifdef LB2_PLAN_TIMING
#include <cstdio>
#endif

namespace lb2 {

constexpr int PLAN_THREADS = 128;

// One build-attribute note.  16-byte halves: {start,end} and the packed rest, so a merge step loads
// two vectors per element.
struct __align__(16) DNote {
  uint64_t start, end;
  uint16_t type;    // 0x100 OPEN, 0x101 FUNC, 0 deleted
  uint16_t nrank;   // rank of the attribute name among the distinct names (valid when !names_ambiguous)
  uint16_t off;     // offset of the note header inside the section
  uint16_t namesz;
  uint16_t cls;     // index of the first note with the identical name (equality class)
  uint8_t ver;      // is a "GA$<version>" note
  uint8_t pad;
  uint32_t tag;     // name hash and length packed: one compare filters class candidates
};
static_assert(sizeof(DNote) == 32, "DNote is read as two 16-byte vectors");

struct NoteWork {
  __align__(16) uint8_t buf[MAX_NOTE_BYTES];  // the section's bytes: a dependent load from here costs ~30 cycles, from L2 ~300
  DNote notes[MAX_NOTES];
  uint64_t key[MAX_NOTES];   // bytes 3..10 of the name, big-endian packed, zero padded: decides most name comparisons
  uint16_t perm[MAX_NOTES], tmp[MAX_NOTES];
};
struct ExtWork {
  uint64_t src[MAX_EXT], dst[MAX_EXT], len[MAX_EXT];
  uint32_t tiles[MAX_EXT];
};

struct PlanSmem {
  Ehdr eh;
  Shdr sh[MAX_SH];
  Phdr ph[MAX_PH];
  Phdr nph[MAX_PH];
  uint64_t new_off[MAX_SH];
  uint64_t new_size[MAX_SH];
  uint64_t src_addr[MAX_SH];  // absolute device address of the section's bytes (input or scratch)
  uint64_t memb[MAX_PH];      // kept sections program header j carries (BFD ELF_SECTION_IN_SEGMENT)
  uint64_t seg_mask[MAX_PH];  // kept alloc sections laid out with PT_LOAD j
  uint64_t seg_bits[MAX_PH];  // ... of which have file contents (not NOBITS)
  uint64_t load_rel_end[MAX_PH], load_mem_top[MAX_PH], load_newoff[MAX_PH], load_base[MAX_PH];
  uint64_t ent_key[MAX_SH + 1];   // last 8 characters of each unique name, reversed
  uint32_t ent_off[MAX_SH + 1];
  uint16_t ent_str[MAX_SH + 1];
  uint16_t ent_len[MAX_SH + 1];
  int16_t ent_host[MAX_SH + 1];
  uint8_t ent_sorted[MAX_SH + 1];
  uint8_t ent_pos[MAX_SH + 1];    // position of each entry in the sorted order
  uint8_t sec_ent[MAX_SH];
  int8_t seg[MAX_SH];
  uint8_t new_index[MAX_SH];
  uint8_t order[MAX_SH + 1];
  uint8_t pkeep[MAX_PH];
  uint8_t piece[MAX_SH + 1];
  uint32_t ord_hash[MAX_SH + 1]; // name hash / length per OUTPUT position: the name searches below read them in a row
  uint16_t ord_len[MAX_SH + 1];
  uint64_t pos_off[MAX_SH + 1];  // new offset of the content section at each output position, ~0 for the others (phase L)
  uint16_t name_len[MAX_SH];     // strlen of each section's name
  uint32_t name_hash[MAX_SH];    // FNV-1a of each section's name: cheap inequality test
  char names[MAX_STR + 48];      // + ".shstrtab" literal + slack for the 24-byte name loads
  uint8_t name_kind[MAX_SH];     // 1: the section is called .dynstr, 2: .dynsym (what sh_link must point at)
  // build-attribute note workspace (phase E, warp 1) and the extent list (phases L/M) are never live together
  union { NoteWork n; ExtWork x; } u;
  // scalars shared by the CTA
  int nflag[6];                 // note merging: corrupt, v1, v2 / ambiguous names, v3, not-an-order, walk verdict
  int n_notes;
  uint32_t n_newsize;
  uint64_t defer_mask;          // note sections with more than 32 notes: merged by the whole CTA after the join
  uint32_t defer_scr_used;
  int fail, fail_e, err_mal, err_uns, need_hoist, names_ambiguous;
  int nk, nent, n_ext, new_phnum;
  uint32_t keep32[2], alloc32[2], nobits32[2];
  uint32_t strsz, new_strsz;
  uint64_t cur, shstr_off, new_shoff, total, hdr_bytes, note_hdr_bytes, copy_bytes;
  unsigned long long tile_base;
  unsigned long long big_ext[3];  // bit e: extent e has more than 64 tiles (MAX_EXT = 140 extents)
  unsigned long long vbig_ext[3]; // ... more than BIG_EXT_TILES tiles
  unsigned long long mid_ext[3];  // bit e: extent e has 5..64 tiles
  int vbig_taken;                 // the expand kernel's list had room: the CTA does not write those tiles itself
  uint32_t n_tiles;
#ifdef LB2_PLAN_TIMING
  long long t_warp[4];
#endif
};

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ uint64_t lowbit(uint64_t v) { return v & (~v + 1); }
__device__ __forceinline__ uint64_t align_up(uint64_t v, uint64_t a) { return a > 1 ? (v + a - 1) / a * a : v; }

__device__ __forceinline__ int d_strlen(const char *s) { int n = 0; while (s[n]) n++; return n; }
__device__ __forceinline__ bool d_streq(const char *a, const char *b) {
  for (int i = 0;; i++) { if (a[i] != b[i]) return false; if (!a[i]) return true; }
}
__device__ __forceinline__ bool d_prefix(const char *s, const char *p) {
  for (int i = 0; p[i]; i++) if (s[i] != p[i]) return false;
  return true;
}

__device__ __forceinline__ uint32_t d_hash(const char *s, int *len_out) {
  uint32_t h = 2166136261u;
  int n = 0;
  for (; s[n]; n++) h = (h ^ (uint8_t)s[n]) * 16777619u;
  *len_out = n;
  return h;
}

// ---- section names as packed words ------------------------------------------------------------------
// Every rule on section names below is "equals L", "starts with L" or "is L or L.<suffix>" for a literal L
// of at most 21 characters.  A thread packs the first 24 bytes of its section's name into three 64-bit
// words once (bytes behind the terminating NUL forced to zero); each rule is then one or two masked integer
// compares against compile-time constants -- no per-character loops, no divergence between the threads of
// a warp that look at different names.  (Character loops cost ~5 cycles per dependent instruction on a
// latency-bound warp: the string-compare version of this gate was a third of the kernel's critical path.)
struct NameW { uint64_t w[3]; };
struct Lit { uint64_t w[3]; int len; };
template <int N> __host__ __device__ constexpr Lit lit(const char (&s)[N]) {
  Lit l{{0, 0, 0}, N - 1};
  for (int i = 0; i < N - 1; i++) l.w[i >> 3] |= (uint64_t)(uint8_t)s[i] << (8 * (i & 7));
  return l;
}
__device__ __forceinline__ void name_words(const char *s, NameW &n) {  // s has >= 24 readable bytes
  uint64_t w[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    uint64_t v = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) v |= (uint64_t)(uint8_t)s[8 * k + q] << (8 * q);
    w[k] = v;
  }
  bool ended = false;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (ended) { w[k] = 0; continue; }
    const uint64_t z = (w[k] - 0x0101010101010101ull) & ~w[k] & 0x8080808080808080ull;  // lowest set bit: first zero byte
    if (z) {
      const int idx = (__ffsll((long long)z) - 1) >> 3;
      w[k] &= idx ? (~0ull >> (8 * (8 - idx))) : 0ull;
      ended = true;
    }
  }
  n.w[0] = w[0]; n.w[1] = w[1]; n.w[2] = w[2];
}
__device__ __forceinline__ bool nm_pfx(const NameW &n, const Lit &l) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int cnt = l.len - 8 * k;
    if (cnt <= 0) continue;
    const uint64_t m = cnt >= 8 ? ~0ull : (~0ull >> (8 * (8 - cnt)));
    ok = ok && ((n.w[k] & m) == l.w[k]);
  }
  return ok;
}
__device__ __forceinline__ unsigned nm_byte(const NameW &n, int i) { return (unsigned)(n.w[i >> 3] >> (8 * (i & 7))) & 0xffu; }
__device__ __forceinline__ bool nm_eq(const NameW &n, const Lit &l) { return nm_pfx(n, l) && nm_byte(n, l.len) == 0; }
__device__ __forceinline__ bool nm_is(const NameW &n, const Lit &l) {  // "base" or "base.*"
  const unsigned c = nm_byte(n, l.len);
  return nm_pfx(n, l) && (c == 0 || c == '.');
}
#define LB2_L(s) lit(s)

// R1: BFD marks these non-alloc names SEC_DEBUGGING; strip removes them.
__device__ __forceinline__ bool is_debug_name(const NameW &n) {
  return nm_pfx(n, LB2_L(".debug")) || nm_pfx(n, LB2_L(".zdebug")) || nm_pfx(n, LB2_L(".gnu.debuglto_.debug_")) ||
         nm_pfx(n, LB2_L(".gnu.linkonce.wi.")) || nm_pfx(n, LB2_L(".line")) || nm_pfx(n, LB2_L(".stab")) || nm_eq(n, LB2_L(".gdb_index"));
}

// ---- conservative input gate ------------------------------------------------------------------
// BFD normalises several header fields from its own tables (section type and flags by NAME, sh_link by
// looking up .dynstr/.dynsym, LMA from p_paddr ...).  Files written by ld, gold, lld, patchelf or
// objcopy already hold BFD's values; anything else is reported LB2_ST_UNSUPPORTED_LAYOUT (-> host
// strip) rather than guessed at.  Mirrors the gate of the test oracle; found by structure fuzzing.
__device__ __forceinline__ int expected_type_by_name(const NameW &n) {  // bfd/elf.c special_sections_*; -1: not a special name
  // (rules for different second characters exclude each other; inside one letter the order is BFD's)
  if (nm_is(n, LB2_L(".bss"))) return SHT_NOBITS;
  if (nm_eq(n, LB2_L(".comment"))) return SHT_PROGBITS;
  if (nm_is(n, LB2_L(".data")) || nm_is(n, LB2_L(".data1")) || nm_pfx(n, LB2_L(".debug"))) return SHT_PROGBITS;
  if (nm_eq(n, LB2_L(".dynamic"))) return SHT_DYNAMIC;
  if (nm_eq(n, LB2_L(".dynstr"))) return SHT_STRTAB;
  if (nm_eq(n, LB2_L(".dynsym"))) return SHT_DYNSYM;
  if (nm_eq(n, LB2_L(".fini"))) return SHT_PROGBITS;
  if (nm_is(n, LB2_L(".fini_array"))) return SHT_FINI_ARRAY;
  if (nm_eq(n, LB2_L(".got"))) return SHT_PROGBITS;
  if (nm_eq(n, LB2_L(".gnu.version"))) return (int)SHT_GNU_VERSYM;
  if (nm_eq(n, LB2_L(".gnu.version_d"))) return (int)SHT_GNU_VERDEF;
  if (nm_eq(n, LB2_L(".gnu.version_r"))) return (int)SHT_GNU_VERNEED;
  if (nm_eq(n, LB2_L(".gnu.hash"))) return (int)SHT_GNU_HASH;
  if (nm_pfx(n, LB2_L(".gnu.linkonce.b"))) return SHT_NOBITS;
  if (nm_pfx(n, LB2_L(".gnu.linkonce.wi."))) return SHT_PROGBITS;
  if (nm_eq(n, LB2_L(".hash"))) return SHT_HASH;
  if (nm_eq(n, LB2_L(".init")) || nm_eq(n, LB2_L(".interp"))) return SHT_PROGBITS;
  if (nm_is(n, LB2_L(".init_array"))) return SHT_INIT_ARRAY;
  if (nm_pfx(n, LB2_L(".line")) || nm_is(n, LB2_L(".ldata")) || nm_is(n, LB2_L(".lrodata"))) return SHT_PROGBITS;
  if (nm_is(n, LB2_L(".lbss"))) return SHT_NOBITS;
  if (nm_pfx(n, LB2_L(".note"))) return SHT_NOTE;
  if (nm_is(n, LB2_L(".noinit"))) return SHT_NOBITS;
  if (nm_eq(n, LB2_L(".plt")) || nm_is(n, LB2_L(".persistent"))) return SHT_PROGBITS;
  if (nm_is(n, LB2_L(".preinit_array"))) return SHT_PREINIT_ARRAY;
  if (nm_is(n, LB2_L(".rodata")) || nm_is(n, LB2_L(".rodata1"))) return SHT_PROGBITS;
  if (nm_pfx(n, LB2_L(".rela"))) return SHT_RELA;
  if (nm_is(n, LB2_L(".rel"))) return SHT_REL;
  if (nm_is(n, LB2_L(".sbss"))) return SHT_NOBITS;
  if (nm_is(n, LB2_L(".sdata"))) return SHT_PROGBITS;
  if (nm_eq(n, LB2_L(".strtab")) || nm_eq(n, LB2_L(".shstrtab"))) return SHT_STRTAB;
  if (nm_eq(n, LB2_L(".symtab"))) return SHT_SYMTAB;
  if (nm_is(n, LB2_L(".tbss"))) return SHT_NOBITS;
  if (nm_is(n, LB2_L(".tdata")) || nm_is(n, LB2_L(".text"))) return SHT_PROGBITS;
  return -1;
}
__device__ bool type_is_known(uint32_t t) {
  switch (t) {
    case SHT_PROGBITS: case SHT_SYMTAB: case SHT_STRTAB: case SHT_RELA: case SHT_HASH: case SHT_DYNAMIC: case SHT_NOTE: case SHT_NOBITS:
    case SHT_DYNSYM: case SHT_INIT_ARRAY: case SHT_FINI_ARRAY: case SHT_PREINIT_ARRAY: case 19 /* SHT_RELR */: case SHT_GNU_HASH:
    case SHT_GNU_VERDEF: case SHT_GNU_VERNEED: case SHT_GNU_VERSYM: case 0x70000001u /* SHT_X86_64_UNWIND */: return true;
    default: return false;
  }
}

// BFD ELF_SECTION_IN_SEGMENT (check_vma, !strict): which sections a program header carries.
__device__ bool sec_in_seg(const Shdr &s, const Phdr &p) {
  const uint32_t t = p.p_type;
  const bool tls = (s.sh_flags & SHF_TLS) != 0, alloc = (s.sh_flags & SHF_ALLOC) != 0;
  const uint64_t sz = (tls && s.sh_type == SHT_NOBITS && t != PT_TLS) ? 0 : s.sh_size;
  if (tls) { if (!(t == PT_TLS || t == PT_GNU_RELRO || t == PT_LOAD)) return false; }
  else if (t == PT_TLS || t == PT_PHDR) return false;
  if (!alloc && (t == PT_LOAD || t == PT_DYNAMIC || t == PT_GNU_EH_FRAME || t == PT_GNU_STACK || t == PT_GNU_RELRO ||
                 t == PT_GNU_SFRAME || (t >= PT_GNU_MBIND_LO && t <= PT_GNU_MBIND_HI)))
    return false;
  if (s.sh_type != SHT_NOBITS) {
    if (s.sh_offset < p.p_offset) return false;
    if (s.sh_offset - p.p_offset + sz > p.p_filesz) return false;
  }
  if (alloc) {
    if (s.sh_addr < p.p_vaddr) return false;
    if (s.sh_addr - p.p_vaddr + sz > p.p_memsz) return false;
  }
  if ((t == PT_DYNAMIC || t == PT_NOTE) && s.sh_size == 0 && p.p_memsz != 0) {
    bool ok_off = s.sh_type == SHT_NOBITS || (s.sh_offset > p.p_offset && s.sh_offset - p.p_offset < p.p_filesz);
    bool ok_vma = !alloc || (s.sh_addr > p.p_vaddr && s.sh_addr - p.p_vaddr < p.p_memsz);
    if (!(ok_off && ok_vma)) return false;
  }
  return true;
}

// elf-strtab.c strrevcmp on two names held in sm.names
__device__ int strrev_cmp(const char *a, int la, const char *b, int lb) {
  int l = la < lb ? la : lb;
  const unsigned char *s = reinterpret_cast<const unsigned char *>(a) + la - 1;
  const unsigned char *t = reinterpret_cast<const unsigned char *>(b) + lb - 1;
  while (l--) {
    if (*s != *t) return (int)*s - (int)*t;
    s--, t--;
  }
  return la - lb;
}

// ---------------------------------------------------------------- R9: objcopy merge_gnu_build_notes
// note records are read from the shared-memory copy of the section; offsets stay multiples of 4
__device__ __forceinline__ uint32_t ldg32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }
// Warp-cooperative global->shared copy (16-byte vectors when the source allows it, bytes otherwise).
__device__ __forceinline__ void warp_g2s(void *dst_s, const uint8_t *src_g, uint32_t nbytes, int lane) {
  uint8_t *d = static_cast<uint8_t *>(dst_s);
  if ((reinterpret_cast<uintptr_t>(src_g) & 15) == 0) {
    const uint32_t nv = nbytes >> 4;
    for (uint32_t i = lane; i < nv; i += 32) reinterpret_cast<uint4 *>(d)[i] = __ldg(reinterpret_cast<const uint4 *>(src_g) + i);
    for (uint32_t i = (nv << 4) + lane; i < nbytes; i += 32) d[i] = __ldg(src_g + i);
  } else if ((reinterpret_cast<uintptr_t>(src_g) & 3) == 0) {
    const uint32_t nv = nbytes >> 2;
    for (uint32_t i = lane; i < nv; i += 32) reinterpret_cast<uint32_t *>(d)[i] = __ldg(reinterpret_cast<const uint32_t *>(src_g) + i);
    for (uint32_t i = (nv << 2) + lane; i < nbytes; i += 32) d[i] = __ldg(src_g + i);
  } else {
    for (uint32_t i = lane; i < nbytes; i += 32) d[i] = __ldg(src_g + i);
  }
}
__device__ __forceinline__ void wr32(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
__device__ __forceinline__ void wr64(uint8_t *p, uint64_t v) { wr32(p, (uint32_t)v); wr32(p + 4, (uint32_t)(v >> 32)); }

// memcmp(name1 + 3, name2 + 3, min(namesz) - 3) of objcopy's compare_gnu_build_notes, answered from
// the equality class and the 8-byte key whenever they decide it
__device__ __forceinline__ int cmp_note_names(const PlanSmem &sm, const uint8_t *nbuf, const DNote &a, uint64_t ka, const DNote &b, uint64_t kb) {
  if (a.cls == b.cls) return 0;
  const int m = (int)(a.namesz < b.namesz ? a.namesz : b.namesz) - 3;
  if (m <= 0) return 0;
  if (m >= 8) {
    if (ka != kb) return ka < kb ? -1 : 1;
    const uint8_t *n1 = nbuf + a.off + 12 + 3, *n2 = nbuf + b.off + 12 + 3;
    for (int i = 8; i < m; i++) { const int x = n1[i], y = n2[i]; if (x != y) return x - y; }
    return 0;
  }
  const uint64_t x = ka >> (8 * (8 - m)), y = kb >> (8 * (8 - m));
  return x == y ? 0 : (x < y ? -1 : 1);
}
// first sort: by attribute name, then by range (objcopy.c compare_gnu_build_notes).  FAST: the names of
// distinct classes are totally ordered (no name is a prefix of another), so their ranks decide.
template <bool FAST>
__device__ __forceinline__ int cmp_by_attr(const PlanSmem &sm, const uint8_t *nbuf, const DNote &a, uint16_t pa, const DNote &b, uint16_t pb) {
  if (FAST) {
    if (a.nrank != b.nrank) return a.nrank < b.nrank ? -1 : 1;
  } else {
    const int c = cmp_note_names(sm, nbuf, a, sm.u.n.key[pa], b, sm.u.n.key[pb]);
    if (c) return c;
  }
  if (a.end < b.start) return -1;
  if (a.start > b.end) return 1;
  if (a.start < b.start) return -1;
  if (a.end > b.end) return 1;
  if (a.end < b.end) return -1;
  if (a.type == 0x100 && b.type != 0x100) return -1;
  if (a.type != 0x100 && b.type == 0x100) return 1;
  return 0;
}
// The note routines below run either on ONE warp (NT = 32: warp 1, next to the program-header and .shstrtab jobs --
// the common case of a few notes) or on the WHOLE CTA (NT = 128: after the join, for sections with more than 32
// notes, where the per-note and per-pair loops are worth spreading over four warps).  `t` is the thread's index in
// the group; flags travel through shared memory; group_sync is __syncwarp or __syncthreads.
template <int NT> __device__ __forceinline__ void group_sync() {
  if (NT == 32) __syncwarp(); else __syncthreads();
}

// objcopy sorts the notes with libc qsort(); its first comparator is not antisymmetric for nested
// ranges, so the result depends on the exact comparison sequence.  This image's glibc 2.39 qsort
// is the classic top-down merge sort (msort.c: n1 = n / 2, sort both halves, merge taking the left
// element while cmp(left, right) <= 0).  The recursion tree is restated level by level: at depth d the
// segment of node k is found by halving [0, n) along the bits of k, all merges of one depth are
// independent and run on different threads, deepest level first -- the same comparisons in the same
// order inside every merge as the recursive routine, 2n instead of n log n merge steps deep.  Inside a
// merge the heads of both runs live in registers and the element behind each head is already on its way
// from shared memory while the heads are compared.
__device__ __forceinline__ void msort_node(int n, int depth, int k, int *lo, int *hi) {
  int l = 0, h = n;
  for (int b = depth - 1; b >= 0; b--) {
    const int mid = l + (h - l) / 2;
    if ((k >> b) & 1) l = mid; else h = mid;
  }
  *lo = l; *hi = h;
}
template <int NT, bool FAST>
__device__ void group_msort_notes(PlanSmem &sm, const uint8_t *nbuf, int n, int t) {
  const DNote *__restrict__ notes = sm.u.n.notes;
  uint16_t *__restrict__ perm = sm.u.n.perm;
  uint16_t *__restrict__ tmp = sm.u.n.tmp;
  int depth = 0;
  while ((1 << depth) < n) depth++;
  for (int d = depth - 1; d >= 0; d--) {
    for (int k = t; k < (1 << d); k += NT) {
      int lo, hi;
      msort_node(n, d, k, &lo, &hi);
      const int n1 = (hi - lo) / 2;
      int i = lo, j = lo + n1, w = lo, r1 = n1, r2 = (hi - lo) - n1;
      if (r1 == 0 || r2 == 0) continue;
      uint16_t pa = perm[i], pb = perm[j];
      uint16_t pa_n = r1 > 1 ? perm[i + 1] : pa, pb_n = r2 > 1 ? perm[j + 1] : pb;
      DNote a = notes[pa], b = notes[pb];
      DNote a_n = notes[pa_n], b_n = notes[pb_n];
      while (true) {
        const int c = cmp_by_attr<FAST>(sm, nbuf, a, pa, b, pb);
        if (c <= 0) {
          tmp[w++] = pa; i++;
          if (--r1 == 0) break;
          pa = pa_n; a = a_n;
          if (r1 > 1) { pa_n = perm[i + 1]; a_n = notes[pa_n]; }
        } else {
          tmp[w++] = pb; j++;
          if (--r2 == 0) break;
          pb = pb_n; b = b_n;
          if (r2 > 1) { pb_n = perm[j + 1]; b_n = notes[pb_n]; }
        }
      }
      while (r1 > 0) { tmp[w++] = perm[i++]; r1--; }
      for (int q = lo; q < w; q++) perm[q] = tmp[q];  // the tail of the right run is already in place
    }
    group_sync<NT>();
  }
}

// Wherever the comparator IS a strict weak order, glibc's merge sort is simply a stable sort and the result
// can be written down directly: position = number of elements that sort before (ties by current position).
// All pairs are independent -- no sequential merge steps at all.
//   FIRST = false: objcopy.c sort_gnu_build_notes -- OPEN notes first, start ascending, larger ranges first.
//                  A lexicographic order: always consistent.
//   FIRST = true : objcopy.c compare_gnu_build_notes -- by attribute name, then by range.  Within one name it
//                  equals (start, end, OPEN first) exactly when no note starts inside an earlier-starting note
//                  of the same name without also ending behind it, and no note has start > end (see the case
//                  analysis in DESIGN.md).  The pass checks that while it counts; if it fails, nothing has been
//                  written to perm[] and the caller runs the restated merge sort instead.
template <int NT, bool FIRST>
__device__ void group_ranksort_notes(PlanSmem &sm, int n, int t) {
  DNote *__restrict__ notes = sm.u.n.notes;
  uint16_t *__restrict__ perm = sm.u.n.perm;
  uint16_t *__restrict__ tmp = sm.u.n.tmp;
  for (int p = t; p < n; p += NT) {
    int viol = 0;
    const uint16_t ip = perm[p];
    const DNote a = notes[ip];
    const bool a_open = a.type == 0x100;
    if (FIRST && a.start > a.end) viol = 1;
    int r = 0;
#pragma unroll 2
    for (int q = 0; q < n; q++) {
      const DNote b = notes[perm[q]];
      const bool b_open = b.type == 0x100;
      int c;  // sign of compare(b, a)
      if (FIRST) {
        if (b.nrank != a.nrank) c = b.nrank < a.nrank ? -1 : 1;
        else {
          if (b.start < a.start && a.start <= b.end && a.end <= b.end) viol = 1;
          c = b.start != a.start ? (b.start < a.start ? -1 : 1) : b.end != a.end ? (b.end < a.end ? -1 : 1) : (b_open == a_open ? 0 : (b_open ? -1 : 1));
        }
      } else {
        c = b_open != a_open ? (b_open ? -1 : 1) : b.start != a.start ? (b.start < a.start ? -1 : 1) : b.end != a.end ? (b.end > a.end ? -1 : 1) : 0;
      }
      r += (c < 0) || (c == 0 && q < p);
    }
    tmp[r] = ip;
    if (FIRST && viol) notes[a.cls].pad = 1;   // this attribute's notes are not ordered by the comparator: see below
  }
  group_sync<NT>();
  for (int p = t; p < n; p += NT) perm[p] = tmp[p];
  group_sync<NT>();
}

// The first sort when the names ARE totally ordered but some attribute has nested ranges.  Comparisons between
// notes of different attributes are consistent, so every merge of the library sort interleaves two runs attribute
// by attribute: the final order is the attributes in rank order, and inside one attribute exactly what the same
// merge sort does to that attribute's notes alone -- split points taken from the GLOBAL recursion tree ([lo,hi) ->
// [lo,mid), [mid,hi) over the original indices), merges taking the left element while cmp <= 0.  The rank sort
// has already put every attribute's notes into one contiguous stretch of perm[] (right for the attributes whose
// comparator is an order); for the others one thread per attribute replays the restricted merge sort: O(m log n)
// sequential steps for m notes instead of 2n for the whole section, all such attributes at once.
template <int NT>
__device__ void group_fix_nested_classes(PlanSmem &sm, const uint8_t *nbuf, int n, int t) {
  const DNote *__restrict__ notes = sm.u.n.notes;
  uint16_t *__restrict__ perm = sm.u.n.perm;
  uint16_t *__restrict__ tmp = sm.u.n.tmp;
  // each replay is a private sequential loop: threads of one warp that run different replays are serialised by the
  // hardware, so neighbouring representatives (the first notes of a section, typically) go to different warps
  for (int rep = (NT == 32) ? t : ((t & 31) * (NT / 32) + (t >> 5)); rep < n; rep += NT) {
    if (notes[rep].cls != rep || !notes[rep].pad) continue;
    // the attribute's stretch of perm[]
    int p0 = 0;
    while (p0 < n && notes[perm[p0]].cls != rep) p0++;
    int p1 = p0;
    while (p1 < n && notes[perm[p1]].cls == rep) p1++;
    // back to original order (the values ARE the original indices): insertion sort
    for (int i = p0 + 1; i < p1; i++) {
      const uint16_t v = perm[i];
      int j = i - 1;
      while (j >= p0 && perm[j] > v) { perm[j + 1] = perm[j]; j--; }
      perm[j + 1] = v;
    }
    // msort_with_tmp over the global index tree, restricted to this stretch; explicit stack (depth <= log2(MAX_NOTES) + 2)
    int f_lo[12], f_hi[12], f_a[12], f_b[12], f_s[12], f_stage[12];
    int sp = 0;
    f_lo[0] = 0; f_hi[0] = n; f_a[0] = p0; f_b[0] = p1; f_s[0] = 0; f_stage[0] = 0;
    while (sp >= 0) {
      const int lo = f_lo[sp], hi = f_hi[sp], a = f_a[sp], b = f_b[sp];
      if (b - a <= 1 || hi - lo <= 1) { sp--; continue; }
      const int mid = lo + (hi - lo) / 2;
      if (f_stage[sp] == 0) {
        int sidx = a;
        while (sidx < b && (int)perm[sidx] < mid) sidx++;   // still in original order: no descendant has run yet
        f_s[sp] = sidx; f_stage[sp] = 1;
        sp++; f_lo[sp] = lo; f_hi[sp] = mid; f_a[sp] = a; f_b[sp] = sidx; f_stage[sp] = 0;
      } else if (f_stage[sp] == 1) {
        f_stage[sp] = 2;
        const int sidx = f_s[sp];
        sp++; f_lo[sp] = mid; f_hi[sp] = hi; f_a[sp] = sidx; f_b[sp] = b; f_stage[sp] = 0;
      } else {
        const int sidx = f_s[sp];
        int i = a, j = sidx, w = a, r1 = sidx - a, r2 = b - sidx;
        if (r1 > 0 && r2 > 0) {
          uint16_t pa = perm[i], pb = perm[j];
          DNote na = notes[pa], nb = notes[pb];
          while (true) {
            const int c = cmp_by_attr<true>(sm, nbuf, na, pa, nb, pb);
            if (c <= 0) {
              tmp[w++] = pa; i++;
              if (--r1 == 0) break;
              pa = perm[i]; na = notes[pa];
            } else {
              tmp[w++] = pb; j++;
              if (--r2 == 0) break;
              pb = perm[j]; nb = notes[pb];
            }
          }
          while (r1 > 0) { tmp[w++] = perm[i++]; r1--; }
          for (int q = a; q < w; q++) perm[q] = tmp[q];  // the tail of the right run is already in place
        }
        sp--;
      }
    }
  }
  group_sync<NT>();
}

// 64-bit value of the nearest lane at or below `lane` whose bit is set in `mask`, else `carry`
__device__ __forceinline__ uint64_t last_set_value(unsigned mask, uint64_t v, uint64_t carry, int lane) {
  const unsigned m = mask & (0xffffffffu >> (31 - lane));
  const int src = m ? 31 - __clz((int)m) : lane;
  const uint64_t got = __shfl_sync(0xffffffffu, v, src);
  return m ? got : carry;
}

#ifdef LB2_HOST_EMULATION   // the CPU emulator reports which sort path a file took, so the tests can insist both are covered
static int lb2_path_counts[4];   // [0] rank sort, [1] merge sort with name ranks, [2] merge sort with full name compares, [3] sections merged by the whole CTA
#define LB2_COUNT(k) do { if (t == 0) __atomic_fetch_add(&lb2_path_counts[k], 1, __ATOMIC_RELAXED); } while (0)
#else
#define LB2_COUNT(k) do { } while (0)
#endif
#ifdef LB2_PLAN_TIMING
#define LB2_NT(k) do { group_sync<NT>(); if (t == 0) nt_[k] = clock64(); } while (0)
#else
#define LB2_NT(k) do { } while (0)
#endif

// Step 1 of the merge, shared by both group sizes: one thread walks the variable-length records (three words each)
// and records where every note starts.  Result in sm.n_notes / sm.nflag[5] (1 corrupt, 2 more notes than MAX_NOTES).
__device__ __forceinline__ void walk_note_records(PlanSmem &sm, const uint8_t *nbuf, uint32_t size) {
  DNote *__restrict__ notes = sm.u.n.notes;
  uint16_t *__restrict__ perm = sm.u.n.perm;
  int n = 0, e1 = 0;
  uint32_t remain = size, p = 0;
  while (remain >= 12) {
    if (n >= MAX_NOTES) { e1 = 2; break; }
    const uint32_t namesz = ldg32(nbuf + p), descsz = ldg32(nbuf + p + 4);
    const uint64_t padded = ((uint64_t)namesz + 3) & ~3ull;   // 64-bit: namesz = 0xffffffff must not wrap to 0
    if (((descsz + 3) & ~3u) != descsz) { e1 = 1; break; }
    if (padded + descsz + 12 > remain) { e1 = 1; break; }
    notes[n].off = (uint16_t)p;
    perm[n] = (uint16_t)n;
    remain -= 12 + padded + descsz;
    p += 12 + padded + descsz;
    n++;
  }
  if (!e1 && remain != 0) e1 = 1;
  sm.n_notes = n;
  sm.nflag[5] = e1;
  sm.nflag[0] = sm.nflag[1] = sm.nflag[2] = sm.nflag[3] = sm.nflag[4] = 0;
}

// Merges the `size` bytes of notes at nbuf (the shared-memory copy of the section, already walked by
// walk_note_records) and writes the result to `out` (global scratch).  Returns the new size; *err: 1 = objcopy
// would report corrupt notes.  Collective over the NT threads of the group.
template <int NT>
__device__ uint32_t merge_build_notes(PlanSmem &sm, const uint8_t *nbuf, uint32_t size, uint8_t *out, int *err, int t) {
  DNote *__restrict__ notes = sm.u.n.notes;
  uint16_t *__restrict__ perm = sm.u.n.perm;
  uint16_t *__restrict__ tmp = sm.u.n.tmp;
  uint64_t *__restrict__ nkey = sm.u.n.key;
  const int lane = t & 31;
#ifdef LB2_PLAN_TIMING
  long long nt_[10];
#endif
  LB2_NT(0);
  const int n = sm.n_notes;
  // 2. every thread decodes its notes: checks, raw range, version class, comparison key and name hash
  for (int i = t; i < n; i += NT) {
    DNote &d = notes[i];
    const uint8_t *h = nbuf + d.off;
    const uint32_t namesz = ldg32(h), descsz = ldg32(h + 4), type = ldg32(h + 8);
    const uint32_t padded = (namesz + 3) & ~3u;
    if (type != 0x100 && type != 0x101) { sm.nflag[0] = 1; continue; }
    if (namesz < 3) { sm.nflag[0] = 1; continue; }  // objcopy accepts 2 and then compares namesz - 3 bytes: treat as corrupt
    const uint8_t *nm = h + 12;
    const uint8_t *dw = h + 12 + padded;
    d.namesz = (uint16_t)namesz;
    d.type = (uint16_t)type;
    d.ver = 0; d.pad = 0; d.nrank = 0;
    const uint8_t c0 = nm[0], c1 = nm[1], c2 = nm[2];
    if (c0 == '$' && c1 == 1 && c2 == '1') sm.nflag[1] = 1;
    else if (namesz > 4 && c0 == 'G' && c1 == 'A' && c2 == '$' && nm[3] == 1) {
      d.ver = 1;
      const uint8_t c4 = nm[4];
      if (c4 == '2') sm.nflag[2] = 1;
      else if (c4 == '3') sm.nflag[3] = 1;
      else { sm.nflag[0] = 1; continue; }
    }
    uint64_t start, end;
    if (descsz == 0) start = end = 0;
    else if (descsz == 4) { start = ldg32(dw); end = ~0ull; }
    else if (descsz == 8) { start = ldg32(dw); end = ldg32(dw + 4); }
    else if (descsz == 16) { start = (uint64_t)ldg32(dw) | ((uint64_t)ldg32(dw + 4) << 32); end = (uint64_t)ldg32(dw + 8) | ((uint64_t)ldg32(dw + 12) << 32); }
    else { sm.nflag[0] = 1; continue; }
    if (start > end) start = end;
    d.start = start;   // raw; ranges inherited from earlier notes are filled in by step 3
    d.end = end;
    if (nm[namesz - 1] != 0) { sm.nflag[0] = 1; continue; }
    uint64_t key = 0;
    uint32_t hsh = 2166136261u;
    for (int q = 0; q < (int)namesz; q++) {
      const uint8_t ch = nm[q];
      hsh = (hsh ^ ch) * 16777619u;
      if (q >= 3 && q < 11) key = (key << 8) | ch;
    }
    if (namesz <= 3) key = 0; else if (namesz < 11) key <<= 8 * (11 - namesz);
    nkey[i] = key;
    d.tag = (hsh << 10) ^ namesz;
  }
  group_sync<NT>();
  {
    if (sm.nflag[0]) { *err = 1; return size; }
    bool a1 = sm.nflag[1] != 0, a2 = sm.nflag[2] != 0, a3 = sm.nflag[3] != 0;
    if (!a1 && !a2 && !a3) a3 = true;  // "version note missing - assuming version 3"
    if ((a1 && a2) || (a1 && a3) || (a2 && a3)) { *err = 1; return size; }
    if (!a3 || size < 12) {            // only v3 notes are merged
      for (uint32_t i = t; i < size; i += NT) out[i] = nbuf[i];
      return size;
    }
  }
  LB2_NT(1);
  // 3. a note without a range inherits the previous OPEN / FUNC note's ("if (start) pos = start; start = pos"):
  //    the nearest earlier note of the same kind with a non-zero raw value, found with ballots, 32 notes a round
  //    (the first warp of the group)
  if (t < 32) {
    uint64_t pos = 0, poe = 0, pfs = 0, pfe = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < n;
      uint64_t s = 0, e = 0;
      bool open = false;
      if (valid) { s = notes[i].start; e = notes[i].end; open = notes[i].type == 0x100; }
      const unsigned m_os = __ballot_sync(0xffffffffu, valid && open && s != 0), m_oe = __ballot_sync(0xffffffffu, valid && open && e != 0);
      const unsigned m_fs = __ballot_sync(0xffffffffu, valid && !open && s != 0), m_fe = __ballot_sync(0xffffffffu, valid && !open && e != 0);
      const uint64_t os = last_set_value(m_os, s, pos, lane), oe = last_set_value(m_oe, e, poe, lane);
      const uint64_t fs = last_set_value(m_fs, s, pfs, lane), fe = last_set_value(m_fe, e, pfe, lane);
      if (valid) { notes[i].start = open ? os : fs; notes[i].end = open ? oe : fe; }
      pos = __shfl_sync(0xffffffffu, os, 31); poe = __shfl_sync(0xffffffffu, oe, 31);
      pfs = __shfl_sync(0xffffffffu, fs, 31); pfe = __shfl_sync(0xffffffffu, fe, 31);
    }
  }
  // 4. equality classes of the names (first note with identical name), one packed tag compare per candidate
  //    (independent of step 3: different fields)
  for (int i = t; i < n; i += NT) {
    DNote &d = notes[i];
    const uint8_t *nm = nbuf + d.off + 12;
    const uint32_t tag = d.tag;
    int cls = i;
    for (int j = 0; j < i; j++) {
      if (notes[j].tag != tag) continue;
      // confirm the (near certain) match word by word: names start at 4-aligned offsets of the staged section,
      // namesz is equal (it is part of the tag), bytes behind namesz are masked off in the last word
      const uint32_t *ow = reinterpret_cast<const uint32_t *>(nbuf + notes[j].off + 12), *mw = reinterpret_cast<const uint32_t *>(nm);
      const int nw = (int)d.namesz >> 2, rem = (int)d.namesz & 3;
      bool same = true;
      for (int q = 0; q < nw; q++) if (ow[q] != mw[q]) { same = false; break; }
      if (same && rem) same = ((ow[nw] ^ mw[nw]) & (0xffffffffu >> (8 * (4 - rem)))) == 0;
      if (same) { cls = j; break; }
    }
    d.cls = (uint16_t)cls;
  }
  group_sync<NT>();
  // 4b. rank of every distinct name among the distinct names; if two distinct names compare equal (one is
  //     a prefix of the other beyond byte 3) ranks cannot stand in for the comparator: slow path
  for (int i = t; i < n; i += NT) {
    const DNote d = notes[i];
    if (d.cls != i) continue;
    const uint64_t ki = nkey[i];
    int less = 0, amb1 = 0;
    for (int j = 0; j < n; j++) {
      if (j == i || notes[j].cls != j) continue;
      const int c = cmp_note_names(sm, nbuf, notes[j], nkey[j], d, ki);
      less += c < 0;
      amb1 |= c == 0;
    }
    notes[i].nrank = (uint16_t)less;
    if (amb1) sm.nflag[1] = 1;   // (nflag[1..3] are free again: version flags were consumed above, behind a group sync)
  }
  group_sync<NT>();
  const bool amb = sm.nflag[1] != 0;
  for (int i = t; i < n; i += NT) if (notes[i].cls != i) notes[i].nrank = notes[notes[i].cls].nrank;
  group_sync<NT>();
  LB2_NT(2);
  if (amb) { LB2_COUNT(2); group_msort_notes<NT, false>(sm, nbuf, n, t); }   // restated glibc merge sort, level by level
  else {
    group_ranksort_notes<NT, true>(sm, n, t);
#ifdef LB2_HOST_EMULATION
    { int any = 0; for (int i = 0; i < n; i++) any |= notes[i].pad; if (any) LB2_COUNT(1); else LB2_COUNT(0); }
#endif
    group_fix_nested_classes<NT>(sm, nbuf, n, t);
  }
  LB2_NT(3);
  // 5. objcopy's merge pass: every note looks back over the SURVIVING notes of the same attribute (at most 17).
  //    The survivors so far are kept as a stack in tmp[], so deleted notes cost nothing to skip.
  if (t == 0) {
    int nl = 0;
    for (int i = 0; i < n; i++) {
      const uint16_t pi = perm[i];
      DNote pn = notes[pi];
      if (pn.type == 0) continue;
      if (pn.start == pn.end) { notes[pi].type = 0; continue; }
      int iter = 0;
      bool dead = false;
      for (int b = nl - 1; b >= 0; b--) {
        DNote &back = notes[tmp[b]];
        const uint64_t bs = back.start, be = back.end;
        if (back.cls != pn.cls) break;  // a different attribute name ends the search
        if (bs == pn.start && be == pn.end) { dead = true; break; }
        if (pn.start >= bs && pn.end <= be) { dead = true; break; }
        bool merge;
        if (be < pn.start) merge = (((be + 15) & ~15ull) < pn.start);
        else merge = (be != pn.end);
        if (back.type != pn.type) merge = false;  // OPEN and FUNC notes are never combined
        if (merge) {
          if (pn.start < bs) back.start = pn.start;
          if (pn.end > be) back.end = pn.end;
          dead = true;
          break;
        }
        if (iter++ > 16) break;
      }
      if (dead) notes[pi].type = 0;
      else tmp[nl++] = pi;
    }
  }
  group_sync<NT>();
  LB2_NT(4);
  group_ranksort_notes<NT, false>(sm, n, t);
  LB2_NT(5);
  // 6. output offsets and range elision: a surviving note drops its description when its range equals the
  //    previous survivor's.  Ballot + shuffle scan, 32 sorted positions a round (the first warp of the group).
  if (t < 32) {
    uint64_t ps = 0, pe = 0;
    uint32_t run = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      DNote pn;
      pn.type = 0; pn.start = pn.end = 0; pn.namesz = 0;
      if (i < n) pn = notes[perm[i]];
      const bool surv = i < n && pn.type != 0;
      const unsigned ms = __ballot_sync(0xffffffffu, surv);
      // previous survivor strictly below this lane, else the carry from earlier rounds
      const unsigned below = lane ? (ms & (0xffffffffu >> (32 - lane))) : 0u;
      const int src = below ? 31 - __clz((int)below) : lane;
      const uint64_t qs = __shfl_sync(0xffffffffu, pn.start, src), qe = __shfl_sync(0xffffffffu, pn.end, src);
      const uint64_t prev_s = below ? qs : ps, prev_e = below ? qe : pe;
      const bool elide = surv && pn.start == prev_s && pn.end == prev_e;
      const uint32_t v = surv ? 12u + ((pn.namesz + 3u) & ~3u) + (elide ? 0u : 16u) : 0u;
      uint32_t inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      if (i < n) tmp[i] = surv ? (uint16_t)(((run + inc - v) >> 2) | (elide ? 0x8000u : 0u)) : (uint16_t)0xffff;  // offsets are multiples of 4, < 64 KB
      run += __shfl_sync(0xffffffffu, inc, 31);
      if (ms) {
        const int last = 31 - __clz((int)ms);
        ps = __shfl_sync(0xffffffffu, pn.start, last); pe = __shfl_sync(0xffffffffu, pn.end, last);
      }
    }
    if (t == 0) sm.n_newsize = run;
  }
  group_sync<NT>();
  const uint32_t newsize = sm.n_newsize;
  if (newsize >= size) {  // objcopy keeps the original contents unless the merged notes are smaller
    for (uint32_t i = t; i < size; i += NT) out[i] = nbuf[i];
    group_sync<NT>();
    return size;
  }
  for (int i = t; i < n; i += NT) {
    const uint16_t tt = tmp[i];
    if (tt == 0xffff) continue;
    const DNote pn = notes[perm[i]];
    const bool elide = (tt & 0x8000u) != 0;
    uint8_t *o = out + ((uint32_t)(tt & 0x7fffu) << 2);
    const uint32_t padded = (pn.namesz + 3u) & ~3u;
    wr32(o, pn.namesz);
    wr32(o + 4, elide ? 0u : 16u);
    wr32(o + 8, pn.type);
    const uint8_t *nm = nbuf + pn.off + 12;
    for (uint32_t q = 0; q < padded; q++) o[12 + q] = q < pn.namesz ? nm[q] : 0;
    if (!elide) { wr64(o + 12 + padded, pn.start); wr64(o + 20 + padded, pn.end); }
  }
  group_sync<NT>();
  LB2_NT(6);
#ifdef LB2_PLAN_TIMING
  if (t == 0 && blockIdx.x == 0) printf("  notes n=%d amb=%d threads=%d: parse=%lld aids=%lld sort1=%lld merge=%lld sort2=%lld out=%lld\n", n, (int)amb, NT, nt_[1]-nt_[0], nt_[2]-nt_[1], nt_[3]-nt_[2], nt_[4]-nt_[3], nt_[5]-nt_[4], nt_[6]-nt_[5]);
#endif
  return newsize;
}

#ifdef LB2_PLAN_TIMING
#define LB2_T(k) do { if (tid == 0) t_[k] = clock64(); } while (0)
#else
#define LB2_T(k) do { } while (0)
#endif

// CTA-cooperative global->shared copy.  16-byte vector loads when both sides allow it (the Shdr table of
// a BFD/ld/lld-written file sits at an 8- or 16-aligned e_shoff and every file base in the arena is
// 16-aligned), 8-byte, then byte loads otherwise.
__device__ __forceinline__ void block_g2s(void *dst_s, const uint8_t *src_g, uint32_t nbytes, int tid) {
  uintptr_t s = reinterpret_cast<uintptr_t>(src_g);
  uint8_t *d = static_cast<uint8_t *>(dst_s);
  if (((s | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
    uint32_t nv = nbytes >> 4;
    for (uint32_t i = tid; i < nv; i += PLAN_THREADS)
      reinterpret_cast<uint4 *>(d)[i] = __ldg(reinterpret_cast<const uint4 *>(src_g) + i);
    for (uint32_t i = (nv << 4) + tid; i < nbytes; i += PLAN_THREADS) d[i] = __ldg(src_g + i);
  } else if (((s | reinterpret_cast<uintptr_t>(d)) & 7) == 0) {
    uint32_t nv = nbytes >> 3;
    for (uint32_t i = tid; i < nv; i += PLAN_THREADS)
      reinterpret_cast<uint2 *>(d)[i] = __ldg(reinterpret_cast<const uint2 *>(src_g) + i);
    for (uint32_t i = (nv << 3) + tid; i < nbytes; i += PLAN_THREADS) d[i] = __ldg(src_g + i);
  } else {
    for (uint32_t i = tid; i < nbytes; i += PLAN_THREADS) d[i] = __ldg(src_g + i);
  }
}

#define LB2_REJECT(code) do { if (tid == 0) { a.status[f] = (code); a.out_size[f] = 0; atomicAdd(&a.ctr->n_unsupported, 1u); } return; } while (0)

__global__ void __launch_bounds__(PLAN_THREADS) lb2_plan_kernel(PlanArgs a) {
  __shared__ PlanSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t f = blockIdx.x;
  if (f >= a.n_files) return;
#ifdef LB2_PLAN_TIMING
  long long t_[16];
  for (int q = 0; q < 16; q++) t_[q] = 0;
#endif
  LB2_T(0);
  const uint64_t base = a.in_off[f];
  const uint64_t n = a.in_size[f];
  const uint8_t *in = a.in + base;
  uint8_t *scr = a.scratch + (uint64_t)f * SCR_STRIDE;

  // ---- A. Ehdr (every exit below is taken by the whole CTA: verdicts come from shared memory)
  if (n < 64) LB2_REJECT(ST_NOT_ELF);
  if (tid < 4) reinterpret_cast<uint4 *>(&sm.eh)[tid] = __ldg(reinterpret_cast<const uint4 *>(in) + tid);
  if (tid == 0) { sm.fail = 0; sm.fail_e = 0; sm.defer_mask = 0; sm.defer_scr_used = 0; sm.err_mal = 0; sm.err_uns = 0; sm.need_hoist = 0; sm.names_ambiguous = 0; sm.note_hdr_bytes = 0; }
  __syncthreads();
  const Ehdr &eh = sm.eh;
  int st = ST_OK;
  if (!(eh.e_ident[0] == 0x7f && eh.e_ident[1] == 'E' && eh.e_ident[2] == 'L' && eh.e_ident[3] == 'F')) st = ST_NOT_ELF;
  else if (eh.e_ident[4] != 2 || eh.e_ident[5] != 1) st = ST_NOT_ELF64LE;
  else if (eh.e_type != 2 && eh.e_type != 3) st = ST_BAD_TYPE;
  else if (eh.e_machine != 62 && eh.e_machine != 183) st = ST_UNSUPPORTED_LAYOUT;
  else if (eh.e_version != 1 || eh.e_ident[6] != 1 || eh.e_ehsize != 64) st = ST_UNSUPPORTED_LAYOUT;  // gate: BFD writes these itself
  else if (eh.e_shoff == 0 || eh.e_shnum == 0) st = ST_NO_SECTIONS;
  else if (eh.e_shentsize != 64 || (eh.e_phnum && eh.e_phentsize != 56)) st = ST_MALFORMED;
  else if (eh.e_shstrndx == 0xffff || eh.e_shnum >= 0xff00 || eh.e_phnum == 0xffff) st = ST_XINDEX;
  else if (eh.e_shoff > n || (uint64_t)eh.e_shnum * 64 > n - eh.e_shoff) st = ST_MALFORMED;
  else if (eh.e_phoff > n || (uint64_t)eh.e_phnum * 56 > n - eh.e_phoff) st = ST_MALFORMED;
  else if (eh.e_shstrndx >= eh.e_shnum) st = ST_MALFORMED;
  else if (eh.e_phnum && eh.e_phoff != 64) st = ST_UNSUPPORTED_LAYOUT;
  else if (eh.e_shnum > MAX_SH || eh.e_phnum > MAX_PH) st = ST_PLANNER_LIMIT;
  if (st != ST_OK) LB2_REJECT(st);
  const int shnum = eh.e_shnum, phnum = eh.e_phnum;

  LB2_T(1);
  // ---- B. section headers, program headers, section names: coalesced vector loads into smem
  block_g2s(sm.sh, in + eh.e_shoff, (uint32_t)shnum * 64, tid);
  if (phnum) block_g2s(sm.ph, in + eh.e_phoff, (uint32_t)phnum * 56, tid);
  __syncthreads();
  {
    const Shdr &strh = sm.sh[eh.e_shstrndx];
    if (strh.sh_type != SHT_STRTAB || strh.sh_offset > n || strh.sh_size > n - strh.sh_offset || strh.sh_size == 0) st = ST_MALFORMED;
    else if (strh.sh_size > MAX_STR) st = ST_PLANNER_LIMIT;
    if (st != ST_OK) LB2_REJECT(st);
    const uint32_t strsz0 = (uint32_t)strh.sh_size;
    // byte loads unless the table happens to be 16-aligned (it rarely is; it is <= 2 KB)
    block_g2s(sm.names, in + strh.sh_offset, strsz0, tid);
    if (tid < 10) sm.names[strsz0 + tid] = ".shstrtab"[tid];  // literal appended behind the table
    if (tid == 0) { sm.strsz = strsz0; sm.hdr_bytes = (uint64_t)shnum * 64 + strsz0 + (uint64_t)phnum * 56 + 64; }
  }
  __syncthreads();
  const uint32_t strsz = sm.strsz;
  if (sm.names[strsz - 1] != 0) LB2_REJECT(ST_MALFORMED);

  LB2_T(2);
  // ---- C. input gate + R1 keep/drop: thread i <-> section i (warps 0-1), thread 64+j <-> program header j (warp 2)
  if (tid >= 64 && tid - 64 < phnum) {
    const int j = tid - 64;
    const Phdr &p = sm.ph[j];
    int err_uns = 0;
    if (p.p_paddr != p.p_vaddr) err_uns = 1;                                   // section LMAs come from p_paddr
    if (p.p_align & (p.p_align - 1)) err_uns = 1;                              // BFD: "invalid alignment"
    if (p.p_type == PT_LOAD) {
      if (p.p_align > 1 && ((p.p_vaddr - p.p_offset) & (p.p_align - 1))) err_uns = 1;
      if (p.p_filesz > p.p_memsz || p.p_vaddr + p.p_memsz < p.p_vaddr || p.p_offset + p.p_filesz < p.p_offset) err_uns = 1;
      for (int l = 0; l < j; l++) {                                            // ascending, non-overlapping LOADs
        const Phdr &o = sm.ph[l];
        if (o.p_type != PT_LOAD) continue;
        if (p.p_vaddr < o.p_vaddr + o.p_memsz) err_uns = 1;
        if (p.p_filesz && o.p_filesz && p.p_offset < o.p_offset + o.p_filesz) err_uns = 1;
      }
    }
    if (p.p_type == PT_PHDR && (p.p_offset != 64 || p.p_filesz != (uint64_t)phnum * 56 || p.p_memsz != (uint64_t)phnum * 56)) err_uns = 1;
    if (p.p_type == PT_GNU_STACK && (p.p_offset || p.p_vaddr || p.p_filesz || p.p_memsz)) err_uns = 1;
    if (err_uns) sm.err_uns = 1;
  }
  if (tid >= 96 && tid < 104 && reinterpret_cast<const uint64_t *>(&sm.sh[0])[tid - 96] != 0) sm.err_uns = 1;  // section 0: the all-zero NULL header
  NameW nw;
  nw.w[0] = nw.w[1] = nw.w[2] = 0;
  bool name_ok = false;
  if (tid < shnum) {  // C0: every section's name -> packed words, hash, length, and whether it is .dynstr / .dynsym
    const Shdr &h = sm.sh[tid];
    uint8_t kind = 0;
    if (h.sh_name < strsz) {
      name_ok = true;
      name_words(sm.names + h.sh_name, nw);
      { int ln; sm.name_hash[tid] = d_hash(sm.names + h.sh_name, &ln); sm.name_len[tid] = (uint16_t)ln; }
      kind = nm_eq(nw, LB2_L(".dynstr")) ? 1 : (nm_eq(nw, LB2_L(".dynsym")) ? 2 : 0);
    }
    sm.name_kind[tid] = kind;
  }
  __syncthreads();
  if (tid < 64) {
    const int i = tid;
    int is_keep = 0, is_alloc = 0, is_nobits = 0;
    if (i < shnum) {
      int err_mal = 0, err_uns = 0;
      Shdr &h = sm.sh[i];
      sm.seg[i] = -1; sm.new_size[i] = h.sh_size; sm.new_off[i] = 0;
      sm.src_addr[i] = reinterpret_cast<uint64_t>(in) + h.sh_offset;
      if (!name_ok) err_mal = 1;
      else if (h.sh_type != SHT_NOBITS && h.sh_type != SHT_NULL && (h.sh_offset > n || h.sh_size > n - h.sh_offset)) err_mal = 1;
      else {
        if (i == 0) is_keep = 1;
        else {
          const bool alloc = (h.sh_flags & SHF_ALLOC) != 0;
          bool drop = false;
          if (h.sh_type == SHT_SYMTAB || h.sh_type == SHT_SYMTAB_SHNDX) drop = true;
          else if (h.sh_type == SHT_STRTAB && !alloc) drop = true;
          else if (!alloc && is_debug_name(nw)) drop = true;
          if (h.sh_type == SHT_NULL || h.sh_type == SHT_GROUP) err_uns = 1;
          if ((h.sh_type == SHT_DYNSYM || h.sh_type == SHT_SYMTAB || h.sh_type == SHT_RELA) && h.sh_entsize != 24) err_uns = 1;
          if (h.sh_type == SHT_GNU_VERSYM && h.sh_entsize != 2) err_uns = 1;
          if (h.sh_type == 19 /* SHT_RELR */ && h.sh_entsize != 8) err_uns = 1;
          if (h.sh_type == SHT_REL && h.sh_entsize != 16) err_uns = 1;
          if (!alloc && (h.sh_type == SHT_REL || h.sh_type == SHT_RELA)) err_uns = 1;
          {  // ---- gate (see expected_type_by_name)
            const uint64_t ALLOWED = 0x1 | 0x2 | 0x4 | 0x10 | 0x20 | 0x40 | 0x400 | 0x800 | 0x200000 | 0x10000000;
            const int want = expected_type_by_name(nw);
            if (h.sh_flags & ~ALLOWED) err_uns = 1;
            if (!type_is_known(h.sh_type)) err_uns = 1;
            if (h.sh_type == SHT_NOBITS && !alloc) err_uns = 1;
            if (want >= 0 && (uint32_t)want != h.sh_type && !(h.sh_type == 0x70000001u && want == (int)SHT_PROGBITS)) err_uns = 1;
            if ((h.sh_flags & SHF_INFO_LINK) && h.sh_type != SHT_RELA && h.sh_type != SHT_REL) err_uns = 1;
            if (h.sh_link >= (uint32_t)shnum) err_uns = 1;
            else {
              const Shdr &lk = sm.sh[h.sh_link];
              const int lkind = sm.name_kind[h.sh_link];   // 1 .dynstr, 2 .dynsym, 0 anything else (or a name out of range)
              switch (h.sh_type) {
                case SHT_DYNSYM: case SHT_DYNAMIC: case SHT_GNU_VERDEF: case SHT_GNU_VERNEED:  // BFD: sh_link := index of .dynstr
                  if (h.sh_link == 0 || lkind != 1) err_uns = 1;
                  if (h.sh_type == SHT_DYNAMIC && h.sh_info != 0) err_uns = 1;
                  if (h.sh_type == SHT_DYNSYM && (h.sh_size % 24 != 0 || h.sh_info > h.sh_size / 24)) err_uns = 1;
                  break;
                case SHT_HASH: case SHT_GNU_HASH: case SHT_GNU_VERSYM:                         // BFD: sh_link := index of .dynsym
                  if (h.sh_link == 0 || lkind != 2 || h.sh_info != 0) err_uns = 1;
                  break;
                case SHT_RELA: case SHT_REL:
                  if (h.sh_link != 0 && lkind != 2) err_uns = 1;
                  if (h.sh_info >= (uint32_t)shnum) err_uns = 1;
                  break;
                case SHT_SYMTAB:  // dropped, but BFD reads it first and refuses a broken one
                  if (h.sh_link == 0 || lk.sh_type != SHT_STRTAB || (lk.sh_flags & SHF_ALLOC)) err_uns = 1;
                  if (h.sh_size % 24 != 0 || h.sh_info > h.sh_size / 24) err_uns = 1;
                  break;
                default:          // ordinary sections (and string tables): BFD writes sh_link = sh_info = 0
                  if (h.sh_link != 0 || h.sh_info != 0) err_uns = 1;
                  break;
              }
            }
          }
          // gate: a TLS NOBITS section (.tbss) must sit where the linker puts it -- at the aligned end of the TLS section
          // before it, or at the start of PT_TLS.  BFD derives the file offset it writes for .tbss from that address;
          // only the natural placement is reproduced (found by header fuzzing: sh_addr moved by 4 bytes -> other sh_offset)
          if (alloc && (h.sh_flags & SHF_TLS) && h.sh_type == SHT_NOBITS) {
            const uint64_t al = h.sh_addralign ? h.sh_addralign : 1;
            if (al & (al - 1)) err_uns = 1;
            int pj = -1;
            for (int j = i - 1; j >= 1; j--)
              if ((sm.sh[j].sh_flags & SHF_TLS) && (sm.sh[j].sh_flags & SHF_ALLOC)) { pj = j; break; }
            if (pj >= 0) {
              const uint64_t pe = sm.sh[pj].sh_addr + sm.sh[pj].sh_size;
              if (h.sh_addr < pe || h.sh_addr - pe >= al || (h.sh_addr & (al - 1))) err_uns = 1;
            } else {
              for (int j = 0; j < phnum; j++)
                if (sm.ph[j].p_type == PT_TLS && h.sh_addr != sm.ph[j].p_vaddr) err_uns = 1;
            }
          }
          is_keep = drop ? 0 : 1;
          is_alloc = alloc;
          is_nobits = h.sh_type == SHT_NOBITS;
        }
      }
      if (err_mal) sm.err_mal = 1;
      if (err_uns) sm.err_uns = 1;
    }
    // the three masks every later phase works from: kept, kept+alloc, kept+NOBITS
    const unsigned mk = __ballot_sync(0xffffffffu, is_keep), ma = __ballot_sync(0xffffffffu, is_keep && is_alloc),
                   mn = __ballot_sync(0xffffffffu, is_keep && is_nobits);
    if (lane == 0) { sm.keep32[warp] = mk; sm.alloc32[warp] = ma; sm.nobits32[warp] = mn; }
  }
  __syncthreads();
  if (sm.err_mal || sm.err_uns) LB2_REJECT(sm.err_mal ? ST_MALFORMED : ST_UNSUPPORTED_LAYOUT);
  const uint64_t keepmask = (uint64_t)sm.keep32[0] | ((uint64_t)sm.keep32[1] << 32);
  const uint64_t allocmask = (uint64_t)sm.alloc32[0] | ((uint64_t)sm.alloc32[1] << 32);
  const uint64_t nobitsmask = (uint64_t)sm.nobits32[0] | ((uint64_t)sm.nobits32[1] << 32);

  LB2_T(3);
  // ---- D. R2 output order (a later dynsym is hoisted in front of the first REL/RELA that uses it) and
  //         new section indices.  Also here, while every thread has its header at hand: BFD keeps a
  //         power-of-two alignment the address honours, min(lowbit(align), lowbit(addr)).
  if (tid < shnum && tid > 0) {
    Shdr &h = sm.sh[tid];
    if (((keepmask >> tid) & 1) && (h.sh_type == SHT_REL || h.sh_type == SHT_RELA) && h.sh_link < (uint32_t)shnum && (int)h.sh_link > tid &&
        ((keepmask >> h.sh_link) & 1) && (sm.sh[h.sh_link].sh_type == SHT_DYNSYM || sm.sh[h.sh_link].sh_type == SHT_SYMTAB))
      sm.need_hoist = 1;
  }
  __syncthreads();
  if (tid < shnum && tid > 0) {
    Shdr &h = sm.sh[tid];
    uint64_t al = h.sh_addralign ? lowbit(h.sh_addralign) : 1;
    if (h.sh_addr) { uint64_t lb = lowbit(h.sh_addr); if (lb < al) al = lb; }
    h.sh_addralign = al;
  }
  if (!sm.need_hoist) {
    if (tid < shnum && ((keepmask >> tid) & 1)) {
      const int k = __popcll(keepmask & ((1ull << tid) - 1));
      sm.new_index[tid] = (uint8_t)k;
      sm.order[k] = (uint8_t)tid;
      sm.ord_hash[k] = sm.name_hash[tid]; sm.ord_len[k] = sm.name_len[tid];
    }
    if (tid == 0) sm.nk = __popcll(keepmask);
  } else if (tid == 0) {
    uint64_t emitted = 0;
    int nk = 0;
    for (int i = 0; i < shnum; i++) {
      if (!((keepmask >> i) & 1) || ((emitted >> i) & 1)) continue;
      const Shdr &h = sm.sh[i];
      if ((h.sh_type == SHT_REL || h.sh_type == SHT_RELA) && h.sh_link < (uint32_t)shnum && (int)h.sh_link > i &&
          ((keepmask >> h.sh_link) & 1) && !((emitted >> h.sh_link) & 1) &&
          (sm.sh[h.sh_link].sh_type == SHT_DYNSYM || sm.sh[h.sh_link].sh_type == SHT_SYMTAB)) {
        sm.order[nk++] = (uint8_t)h.sh_link;
        emitted |= 1ull << h.sh_link;
      }
      sm.order[nk++] = (uint8_t)i;
      emitted |= 1ull << i;
    }
    for (int k = 0; k < nk; k++) {
      sm.new_index[sm.order[k]] = (uint8_t)k;
      sm.ord_hash[k] = sm.name_hash[sm.order[k]]; sm.ord_len[k] = sm.name_len[sm.order[k]];
    }
    sm.nk = nk;
  }
  // ---- F1. which kept sections each program header carries: one header per warp-iteration, two ballots
  for (int j = warp; j < phnum; j += PLAN_THREADS / 32) {
    const Phdr &p = sm.ph[j];
    const bool in0 = lane >= 1 && lane < shnum && ((keepmask >> lane) & 1) && sec_in_seg(sm.sh[lane], p);
    const bool in1 = lane + 32 < shnum && ((keepmask >> (lane + 32)) & 1) && sec_in_seg(sm.sh[lane + 32], p);
    const unsigned m0 = __ballot_sync(0xffffffffu, in0), m1 = __ballot_sync(0xffffffffu, in1);
    if (lane == 0) sm.memb[j] = (uint64_t)m0 | ((uint64_t)m1 << 32);
  }
  __syncthreads();
  const int nk = sm.nk;

  LB2_T(4);
  // =====================================================================================================
  // Three independent jobs on three warps: program headers + LOAD layout | note merging | .shstrtab
  // =====================================================================================================
  if (warp == 0) {
    // ---- F2. the PT_LOAD that lays out each kept alloc section = the first LOAD carrying it
    int err = 0;
    for (int i = lane; i < shnum; i += 32) {
      if (!((allocmask >> i) & 1)) continue;
      int sg = -1;
      for (int j = 0; j < phnum; j++)
        if (sm.ph[j].p_type == PT_LOAD && ((sm.memb[j] >> i) & 1)) { sg = j; break; }
      if (sg < 0) err = 1;
      // gate: file offset and address of a loaded section move together (BFD: "lma adjusted" otherwise)
      else if (sm.sh[i].sh_type != SHT_NOBITS && sm.sh[i].sh_offset - sm.ph[sg].p_offset != sm.sh[i].sh_addr - sm.ph[sg].p_vaddr) err = 1;
      sm.seg[i] = (int8_t)sg;
    }
    if (__ballot_sync(0xffffffffu, err)) { if (lane == 0) sm.fail = ST_UNSUPPORTED_LAYOUT; }
    else {
      // members per LOAD (R11: a LOAD without members, other than the header-bearing one, is deleted)
      int keepj = 1;
      if (lane < phnum) {
        const Phdr &p = sm.ph[lane];
        uint64_t mm = 0;
        if (p.p_type == PT_LOAD) {
          uint64_t earlier = 0;
          for (int l = 0; l < lane; l++) if (sm.ph[l].p_type == PT_LOAD) earlier |= sm.memb[l];
          mm = sm.memb[lane] & allocmask & ~earlier;
        }
        const uint64_t mb = mm & ~nobitsmask;
        sm.seg_mask[lane] = mm;
        sm.seg_bits[lane] = mb;
        if (p.p_type == PT_LOAD && p.p_offset != 0 && !mm) keepj = 0;
        sm.pkeep[lane] = (uint8_t)keepj;
        sm.nph[lane] = p;
        // per-LOAD extents relative to the LOAD's own start: everything the cursor chain below needs
        uint64_t rel_end = 0, mem_top = 0;
        if (p.p_type == PT_LOAD && keepj) {
          if (mb) { const int lb = 63 - __clzll((long long)mb); rel_end = (sm.sh[lb].sh_addr - p.p_vaddr) + sm.sh[lb].sh_size; }
          for (uint64_t q = mm; q; q &= q - 1) {
            const int i = __ffsll((long long)q) - 1;
            const Shdr &h = sm.sh[i];
            if (!(h.sh_type == SHT_NOBITS && (h.sh_flags & SHF_TLS))) { const uint64_t e = h.sh_addr + h.sh_size; if (e > mem_top) mem_top = e; }
          }
        }
        sm.load_rel_end[lane] = rel_end;
        sm.load_mem_top[lane] = mem_top;
      }
      const unsigned km = __ballot_sync(0xffffffffu, lane < phnum && keepj);
      const int new_phnum = __popc(km);
      if (lane == 0) sm.new_phnum = new_phnum;
      __syncwarp();
      // ---- G. R10: the file cursor runs over the LOADs in order (a handful of values per LOAD)
      if (lane == 0) {
        uint64_t cur = 64 + (uint64_t)new_phnum * 56, last_vaddr = 0;
        for (int j = 0; j < phnum; j++) {
          const Phdr &p = sm.ph[j];
          if (p.p_type != PT_LOAD || !sm.pkeep[j]) continue;
          if (p.p_vaddr < last_vaddr) { sm.fail = ST_UNSUPPORTED_LAYOUT; break; }
          last_vaddr = p.p_vaddr;
          const bool first = (p.p_offset == 0);
          const bool contents = sm.seg_bits[j] != 0;
          uint64_t new_off = 0;
          if (!first) { const uint64_t al = p.p_align ? p.p_align : 1; new_off = cur + ((p.p_vaddr - cur) & (al - 1)); }  // p_align is a power of two (gate)
          const uint64_t base_off = first ? cur : new_off;
          const uint64_t file_end = contents ? new_off + sm.load_rel_end[j] : base_off;
          uint64_t mem_end = p.p_vaddr + (first ? cur : 0);
          if (sm.load_mem_top[j] > mem_end) mem_end = sm.load_mem_top[j];
          Phdr &q = sm.nph[j];
          q.p_offset = new_off;
          if (!contents && !first) {
            const uint64_t al = p.p_align > 0x1000 ? p.p_align : 0x1000;
            q.p_offset = cur & (al - 1);
            q.p_filesz = 0;
          } else {
            q.p_filesz = file_end - new_off;
          }
          q.p_memsz = mem_end - p.p_vaddr;
          sm.load_newoff[j] = new_off;
          sm.load_base[j] = base_off;
          if (contents || first) cur = file_end;
        }
        sm.cur = cur;
      }
      __syncwarp();
      // every member section's new offset follows from its LOAD's; the order checks of the sequential
      // walk (a section may not start before the previous one ended) are per-section comparisons
      if (!sm.fail) {
        int bad = 0;
        for (int i = lane; i < shnum; i += 32) {
          if (!((allocmask >> i) & 1)) continue;
          const int j = sm.seg[i];
          if (!sm.pkeep[j]) continue;
          const Phdr &p = sm.ph[j];
          const uint64_t lo = sm.load_newoff[j];
          const Shdr &h = sm.sh[i];
          const uint64_t want = lo + (h.sh_addr - p.p_vaddr);
          const uint64_t before = (1ull << i) - 1;
          const uint64_t prev_bits = sm.seg_bits[j] & before, prev_any = sm.seg_mask[j] & before;
          uint64_t floor_off;  // the cursor of the sequential walk when it reaches this section
          if (prev_bits) { const int pb = 63 - __clzll((long long)prev_bits); floor_off = lo + (sm.sh[pb].sh_addr - p.p_vaddr) + sm.sh[pb].sh_size; }
          else if (prev_any) { const int pf = __ffsll((long long)sm.seg_mask[j]) - 1; floor_off = lo + (sm.sh[pf].sh_addr - p.p_vaddr); }
          else floor_off = (h.sh_type != SHT_NOBITS) ? sm.load_base[j] : want;
          if (h.sh_type != SHT_NOBITS) {
            if (want < floor_off) bad = 1;
            sm.new_off[i] = want;
          } else {
            sm.new_off[i] = floor_off;
          }
        }
        if (__ballot_sync(0xffffffffu, bad)) { if (lane == 0) sm.fail = ST_UNSUPPORTED_LAYOUT; }
      }
      __syncwarp();
      // ---- H. R12: every other program header, one per lane
      int gate_fail = 0;
      if (!sm.fail && lane < phnum && sm.pkeep[lane] && sm.ph[lane].p_type != PT_LOAD) {
        const int j = lane;
        const Phdr &p = sm.ph[j];
        Phdr &q = sm.nph[j];
        const uint32_t t = p.p_type;
        const uint64_t mine = sm.memb[j];
        if (t == PT_PHDR) {
          q.p_filesz = q.p_memsz = (uint64_t)new_phnum * 56;
        } else {
          const int first = mine ? __ffsll((long long)mine) - 1 : -1;
          int last_bits = -1;
          uint64_t aend = 0, fend = 0;
          bool any_alloc = false;
          for (uint64_t mq = mine; mq; mq &= mq - 1) {
            const int i = __ffsll((long long)mq) - 1;
            const Shdr &h = sm.sh[i];
            if (h.sh_type != SHT_NOBITS) { last_bits = i; if (h.sh_offset + h.sh_size > fend) fend = h.sh_offset + h.sh_size; }
            if (h.sh_flags & SHF_ALLOC) { any_alloc = true; if (h.sh_addr + h.sh_size > aend) aend = h.sh_addr + h.sh_size; }
          }
          if (first >= 0 && t != PT_GNU_STACK && t != PT_GNU_RELRO && t != PT_TLS) {
            // gate: a segment that carries sections must describe exactly their extent (BFD recomputes
            // offset / filesz / memsz from the sections; natural files already agree)
            const Shdr &hf = sm.sh[first];
            if (any_alloc && (hf.sh_addr != p.p_vaddr || aend - p.p_vaddr != p.p_memsz)) gate_fail = 1;
            if (hf.sh_type != SHT_NOBITS && hf.sh_offset != p.p_offset) gate_fail = 1;
            if (fend && fend - p.p_offset != p.p_filesz) gate_fail = 1;
          }
          if (t == PT_GNU_STACK) { q.p_offset = 0; q.p_filesz = 0; }
          else if (t == PT_GNU_RELRO) {
            bool ok = false;
            if (first >= 0) {
              const uint64_t start = sm.sh[first].sh_addr, end = start + p.p_memsz;
              for (int l = 0; l < phnum && !ok; l++) {
                if (sm.ph[l].p_type != PT_LOAD || !sm.pkeep[l]) continue;
                const uint64_t lm = sm.seg_mask[l];
                if (!lm) continue;
                const int lf = __ffsll((long long)lm) - 1, ll = 63 - __clzll((long long)lm);
                const Shdr &hl = sm.sh[ll];
                uint64_t lend = hl.sh_addr + ((hl.sh_type == SHT_NOBITS && (hl.sh_flags & SHF_TLS)) ? 0 : hl.sh_size);
                if (!(lend > start && sm.sh[lf].sh_addr < end)) continue;
                for (uint64_t mm = lm; mm; mm &= mm - 1) {
                  const int i = __ffsll((long long)mm) - 1;
                  const Shdr &h = sm.sh[i];
                  if (h.sh_addr >= start && h.sh_addr < end && h.sh_size != 0) {
                    q.p_vaddr = h.sh_addr;
                    q.p_paddr = h.sh_addr + (sm.ph[l].p_paddr - sm.ph[l].p_vaddr);
                    q.p_offset = sm.new_off[i];
                    q.p_memsz = end - q.p_vaddr;
                    q.p_filesz = q.p_memsz;
                    const Phdr &nl = sm.nph[l];
                    if (q.p_filesz > nl.p_vaddr + nl.p_filesz - q.p_vaddr) q.p_filesz = nl.p_vaddr + nl.p_filesz - q.p_vaddr;
                    ok = true;
                    break;
                  }
                }
                break;
              }
            }
            if (!ok) { q.p_type = 0; q.p_flags = 0; q.p_offset = q.p_vaddr = q.p_paddr = q.p_filesz = q.p_memsz = q.p_align = 0; }
          } else if (first < 0) {
            q.p_offset = 0; q.p_filesz = 0; q.p_memsz = 0;
          } else {
            q.p_offset = sm.new_off[first];
            q.p_filesz = 0;
            if (t == PT_TLS) {  // p_memsz := address extent of .tdata/.tbss (gold rounds its value up)
              uint64_t end = p.p_vaddr;
              for (uint64_t mq = mine; mq; mq &= mq - 1) {
                const int i = __ffsll((long long)mq) - 1;
                if (sm.sh[i].sh_addr + sm.sh[i].sh_size > end) end = sm.sh[i].sh_addr + sm.sh[i].sh_size;
              }
              q.p_memsz = end - p.p_vaddr;
            }
            if (last_bits >= 0) {
              q.p_filesz = sm.new_off[last_bits] - q.p_offset + sm.new_size[last_bits];
              if (t == PT_NOTE && (sm.sh[last_bits].sh_flags & SHF_ALLOC)) q.p_memsz = q.p_filesz;
            }
          }
        }
      }
      if (__ballot_sync(0xffffffffu, gate_fail)) { if (lane == 0 && !sm.fail) sm.fail = ST_UNSUPPORTED_LAYOUT; }
    }
#ifdef LB2_PLAN_TIMING
    if (lane == 0) sm.t_warp[0] = clock64();
#endif
  } else if (warp == 1) {
    // ---- E. R9 build-attribute note merging (the new sizes feed the non-alloc layout after the join).
    //         Sections with up to 32 notes are merged right here by this warp; the first bigger one sends itself
    //         and everything behind it to the whole CTA (after the join).
    if (!(a.flags & 1u)) {
      uint32_t scr_used = 0;
      int nfail = 0;
      for (uint64_t mq = keepmask & ~allocmask & ~1ull; mq && !nfail; mq &= mq - 1) {
        const int i = __ffsll((long long)mq) - 1;
        const Shdr &h = sm.sh[i];
        if (h.sh_type != SHT_NOTE) continue;
        if (!d_prefix(sm.names + h.sh_name, ".gnu.build.attributes")) continue;
        if (h.sh_size > (uint64_t)MAX_NOTE_BYTES || scr_used + h.sh_size > MAX_NOTE_BYTES) { nfail = ST_PLANNER_LIMIT; break; }
        int err = 0;
        uint8_t *dst = scr + SCR_NOTES + scr_used;
        warp_g2s(sm.u.n.buf, in + h.sh_offset, (uint32_t)h.sh_size, lane);
        __syncwarp();
        if (lane == 0) walk_note_records(sm, sm.u.n.buf, (uint32_t)h.sh_size);
        __syncwarp();
        if (sm.nflag[5]) { nfail = sm.nflag[5] == 2 ? ST_PLANNER_LIMIT : ST_BAD_NOTES; break; }
        if (sm.n_notes > 32) {
          if (lane == 0) { sm.defer_mask = mq; sm.defer_scr_used = scr_used; }
          break;
        }
        const uint32_t ns = merge_build_notes<32>(sm, sm.u.n.buf, (uint32_t)h.sh_size, dst, &err, lane);
        if (err) { nfail = ST_BAD_NOTES; break; }
        if (lane == 0) {
          sm.new_size[i] = ns;
          sm.src_addr[i] = reinterpret_cast<uint64_t>(dst);
          sm.note_hdr_bytes += h.sh_size;
        }
        scr_used += (ns + 15u) & ~15u;
        __syncwarp();
      }
      if (nfail && lane == 0) sm.fail_e = nfail;  // kept apart from warp 0's verdict: notes are judged first, as in objcopy
    }
#ifdef LB2_PLAN_TIMING
    if (lane == 0) sm.t_warp[1] = clock64();
#endif
  } else if (warp == 2) {
    // ---- J. R6 .shstrtab: unique names (entry 0 = ".shstrtab"), reversed-string rank sort across
    //         lanes, suffix merge, offsets in insertion order.
    for (int k = 1 + lane; k < nk; k += 32) {
      const int i = sm.order[k];
      const char *nm = sm.names + sm.sh[i].sh_name;
      int first = k;
      const uint32_t hh = sm.name_hash[i];
      if (sm.name_len[i] == 9 && d_streq(nm, sm.names + strsz)) first = 0;  // a kept section that is itself called .shstrtab
      else {
        const uint16_t ll = sm.name_len[i];
        for (int q = 1; q < k; q++)
          if (sm.ord_hash[q] == hh && sm.ord_len[q] == ll && d_streq(nm, sm.names + sm.sh[sm.order[q]].sh_name)) { first = q; break; }
      }
      sm.piece[k] = (uint8_t)first;  // order position of the first section with this name
    }
    __syncwarp();
    // entries in insertion order: ".shstrtab", then every first occurrence of a non-empty name.  The entry
    // index is the rank of the owner among owners (ballot + popc), no serial walk.
    {
      int ebase = 1;
      if (lane == 0) { sm.ent_str[0] = (uint16_t)strsz; sm.ent_len[0] = 9; }
      for (int k0 = 0; k0 < nk; k0 += 32) {
        const int k = k0 + lane;
        bool owner = false;
        int i = 0;
        if (k >= 1 && k < nk) { i = sm.order[k]; owner = sm.piece[k] == k && sm.name_len[i] != 0; }
        const unsigned m = __ballot_sync(0xffffffffu, owner);
        if (owner) {
          const int e = ebase + __popc(m & ((1u << lane) - 1));
          sm.ent_str[e] = (uint16_t)sm.sh[i].sh_name;
          sm.ent_len[e] = sm.name_len[i];
          sm.sec_ent[i] = (uint8_t)e;
        }
        ebase += __popc(m);
      }
      if (lane == 0) sm.nent = ebase;
    }
    __syncwarp();
    for (int k = 1 + lane; k < nk; k += 32) {
      const int i = sm.order[k];
      if (sm.name_len[i] == 0) sm.sec_ent[i] = 0xff;                                        // empty name -> sh_name 0
      else if (sm.piece[k] != k) sm.sec_ent[i] = sm.piece[k] == 0 ? 0 : sm.sec_ent[sm.order[sm.piece[k]]];
    }
    const int nent = sm.nent;
    // reversed-suffix keys: the last 8 characters, last one most significant, zero padded -- an integer compare
    // of two keys is elf-strtab.c's strrevcmp whenever one of the names is shorter than 8 or the keys differ
    for (int e = lane; e < nent; e += 32) {
      const char *nm = sm.names + sm.ent_str[e];
      const int len = sm.ent_len[e];
      uint64_t key = 0;
      for (int q = 0; q < 8; q++) key = (key << 8) | (q < len ? (uint8_t)nm[len - 1 - q] : 0);
      sm.ent_key[e] = key;
    }
    __syncwarp();
    for (int e = lane; e < nent; e += 32) {
      const uint64_t ke = sm.ent_key[e];
      const int le = sm.ent_len[e];
      int r = 0;
      for (int o = 0; o < nent; o++) {
        if (o == e) continue;
        const uint64_t ko = sm.ent_key[o];
        bool less;
        if (ko != ke) less = ko < ke;
        else less = strrev_cmp(sm.names + sm.ent_str[o], sm.ent_len[o], sm.names + sm.ent_str[e], le) < 0;  // both >= 8 long
        r += less;
      }
      sm.ent_sorted[r] = (uint8_t)e;  // names are unique => ranks are a permutation
      sm.ent_pos[e] = (uint8_t)r;
    }
    __syncwarp();
    // suffix merge.  BFD walks the sorted array from the end keeping the last name that was not merged
    // ("host") and merges a name that is a suffix of the host.  Sorted by reversed string, a name is a suffix
    // of the host exactly when it is a suffix of its immediate successor (everything between a prefix and a
    // string sorts as an extension of that prefix), so the merged flag is a per-position test and the host is
    // the nearest non-merged position above.
    for (int s0 = 0; s0 < nent; s0 += 32) {
      const int s2 = s0 + lane;
      bool merged = false;
      if (s2 < nent - 1) {
        const int c = sm.ent_sorted[s2], hn = sm.ent_sorted[s2 + 1];
        const int lc = sm.ent_len[c], lh = sm.ent_len[hn];
        const uint64_t kc = sm.ent_key[c], kh = sm.ent_key[hn];
        bool suffix = lh > lc;
        if (suffix) {
          if (lc <= 8) suffix = (lc == 8 ? kh == kc : (lc == 0 ? true : (kh >> (8 * (8 - lc))) == (kc >> (8 * (8 - lc)))));
          else {
            const char *hs = sm.names + sm.ent_str[hn] + (lh - lc), *cs = sm.names + sm.ent_str[c];
            for (int q = 0; q < lc; q++) if (hs[q] != cs[q]) { suffix = false; break; }
          }
        }
        merged = suffix;
      }
      if (s2 < nent) sm.piece[s2] = merged ? 1 : 0;   // (piece[] is free again: reused as the merged flag per sorted position)
    }
    __syncwarp();
    for (int e = lane; e < nent; e += 32) {
      int s2 = sm.ent_pos[e];
      if (!sm.piece[s2]) { sm.ent_host[e] = -1; continue; }
      do { s2++; } while (sm.piece[s2]);   // the last position is never merged
      sm.ent_host[e] = (int16_t)sm.ent_sorted[s2];
    }
    __syncwarp();
    {
      // offsets of the non-merged names: exclusive scan of (len + 1) in insertion order, 32 entries per step
      uint32_t run = 1;
      for (int e0 = 0; e0 < nent; e0 += 32) {
        const int e = e0 + lane;
        const uint32_t v = (e < nent && sm.ent_host[e] < 0) ? sm.ent_len[e] + 1u : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (e < nent && sm.ent_host[e] < 0) sm.ent_off[e] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      __syncwarp();
      for (int e = lane; e < nent; e += 32)
        if (sm.ent_host[e] >= 0) { const int h = sm.ent_host[e]; sm.ent_off[e] = sm.ent_off[h] + (sm.ent_len[h] - sm.ent_len[e]); }
      if (lane == 0) sm.new_strsz = run;
      // blob: zero, then every non-merged name at its offset
      if (run <= MAX_STR + 16) {
        for (uint32_t i = lane; i < ((run + 15u) & ~15u); i += 32) scr[SCR_STR + i] = 0;
        __syncwarp();
        for (int e = lane; e < nent; e += 32)
          if (sm.ent_host[e] < 0) {
            const char *s = sm.names + sm.ent_str[e];
            for (int q = 0; q < sm.ent_len[e]; q++) scr[SCR_STR + sm.ent_off[e] + q] = (uint8_t)s[q];
          }
      }
    }
#ifdef LB2_PLAN_TIMING
    if (lane == 0) sm.t_warp[2] = clock64();
#endif
  }
  __syncthreads();
  if (sm.defer_mask && !sm.fail_e) {
    // ---- E'. note sections with more than 32 notes: the same merge, every per-note and per-pair loop spread over
    //          the four warps (the other jobs are done, the CTA would idle behind warp 1 otherwise)
    uint32_t scr_used = sm.defer_scr_used;
    int nfail = 0;
    for (uint64_t mq = sm.defer_mask; mq && !nfail; mq &= mq - 1) {
      const int i = __ffsll((long long)mq) - 1;
      const Shdr &h = sm.sh[i];
      if (h.sh_type != SHT_NOTE) continue;
      if (!d_prefix(sm.names + h.sh_name, ".gnu.build.attributes")) continue;
      if (h.sh_size > (uint64_t)MAX_NOTE_BYTES || scr_used + h.sh_size > MAX_NOTE_BYTES) { nfail = ST_PLANNER_LIMIT; break; }
      int err = 0;
      uint8_t *dst = scr + SCR_NOTES + scr_used;
      __syncthreads();   // the previous section's workspace is no longer read
      block_g2s(sm.u.n.buf, in + h.sh_offset, (uint32_t)h.sh_size, tid);
      __syncthreads();
      if (tid == 0) walk_note_records(sm, sm.u.n.buf, (uint32_t)h.sh_size);
      __syncthreads();
      if (sm.nflag[5]) { nfail = sm.nflag[5] == 2 ? ST_PLANNER_LIMIT : ST_BAD_NOTES; break; }
#ifdef LB2_HOST_EMULATION
      { const int t = tid; LB2_COUNT(3); }
#endif
      const uint32_t ns = merge_build_notes<PLAN_THREADS>(sm, sm.u.n.buf, (uint32_t)h.sh_size, dst, &err, tid);
      if (err) { nfail = ST_BAD_NOTES; break; }
      if (tid == 0) {
        sm.new_size[i] = ns;
        sm.src_addr[i] = reinterpret_cast<uint64_t>(dst);
        sm.note_hdr_bytes += h.sh_size;
      }
      scr_used += (ns + 15u) & ~15u;
    }
    __syncthreads();
    if (nfail) LB2_REJECT(nfail);
  }
  if (sm.fail_e || sm.fail) LB2_REJECT(sm.fail_e ? sm.fail_e : sm.fail);
  LB2_T(5);

  // ---- I. R4 non-alloc sections packed behind the last allocated byte (align-then-add chain over the
  //         few kept non-alloc sections), R5 .shstrtab and section-table position
  if (tid == 0) {
    uint64_t cur = sm.cur;
    if (!sm.need_hoist) {
      for (uint64_t mq = keepmask & ~allocmask & ~1ull; mq; mq &= mq - 1) {   // index order == output order
        const int i = __ffsll((long long)mq) - 1;
        const Shdr &h = sm.sh[i];
        cur = align_up(cur, h.sh_addralign ? h.sh_addralign : 1);
        sm.new_off[i] = cur;
        if (h.sh_type != SHT_NOBITS) cur += sm.new_size[i];
      }
    } else {
      for (int k = 1; k < nk; k++) {
        const int i = sm.order[k];
        const Shdr &h = sm.sh[i];
        if (h.sh_flags & SHF_ALLOC) continue;
        cur = align_up(cur, h.sh_addralign ? h.sh_addralign : 1);
        sm.new_off[i] = cur;
        if (h.sh_type != SHT_NOBITS) cur += sm.new_size[i];
      }
    }
    sm.cur = cur;
    sm.shstr_off = cur;
    sm.new_shoff = align_up(cur + sm.new_strsz, 8);
    sm.total = sm.new_shoff + (uint64_t)(nk + 1) * 64;
    sm.hdr_bytes += sm.note_hdr_bytes;
  }
  __syncthreads();
  const uint32_t new_strsz = sm.new_strsz;
  const int new_phnum = sm.new_phnum;
  // sanity bound on the stripped size: re-layout can add LOAD alignment padding (at most a few MB per segment),
  // never more; a larger value means wrapped address arithmetic on a hostile or corrupt file.  Also keeps the
  // per-file tile count far inside 32 bits.
  if (sm.total > n + (65ull << 20)) LB2_REJECT(ST_UNSUPPORTED_LAYOUT);
  if (new_strsz > MAX_STR + 16) LB2_REJECT(ST_PLANNER_LIMIT);

  LB2_T(6);
  // ---- K. R7 new section-header table (one header per thread), R8 Ehdr, new Phdr table
  if (tid <= nk) {
    const int k = tid;
    Shdr h;
    if (k == 0) {
      h.sh_name = 0; h.sh_type = 0; h.sh_flags = 0; h.sh_addr = 0; h.sh_offset = 0; h.sh_size = 0; h.sh_link = 0;
      h.sh_info = 0; h.sh_addralign = 0; h.sh_entsize = 0;  // BFD writes a fresh all-zero NULL header
    } else if (k == nk) {
      h.sh_name = sm.ent_off[0]; h.sh_type = SHT_STRTAB; h.sh_flags = 0; h.sh_addr = 0; h.sh_offset = sm.shstr_off;
      h.sh_size = new_strsz; h.sh_link = 0; h.sh_info = 0; h.sh_addralign = 1; h.sh_entsize = 0;
    } else {
      const int i = sm.order[k];
      h = sm.sh[i];
      const char *nm = sm.names + h.sh_name;
      h.sh_name = sm.sec_ent[i] == 0xff ? 0 : sm.ent_off[sm.sec_ent[i]];
      h.sh_offset = sm.new_off[i];
      h.sh_size = sm.new_size[i];
      if (h.sh_link && h.sh_link < (uint32_t)shnum) h.sh_link = ((keepmask >> h.sh_link) & 1) ? sm.new_index[h.sh_link] : 0;
      if ((h.sh_flags & SHF_INFO_LINK) && h.sh_info && h.sh_info < (uint32_t)shnum)
        h.sh_info = ((keepmask >> h.sh_info) & 1) ? sm.new_index[h.sh_info] : 0;
      if ((h.sh_type == SHT_REL || h.sh_type == SHT_RELA) && sm.sh[i].sh_link == 0) {
        // assign_section_numbers(): an allocated reloc section without a symbol table gets .dynsym
        for (int q = 1; q < nk; q++) {
          const int oi = sm.order[q];
          if (sm.name_kind[oi] == 2) { h.sh_link = (uint32_t)q; break; }
        }
      }
      if (h.sh_type == SHT_REL || h.sh_type == SHT_RELA) {
        // BFD re-derives the section a dynamic reloc section applies to from its name
        const char *t = nullptr;
        if (d_prefix(nm, ".rela")) t = nm + 5;
        else if (d_prefix(nm, ".rel")) t = nm + 4;
        int target = -1;
        if (t && *t) {
          auto find_name = [&](const char *want) {
            int wl;
            const uint32_t wh = d_hash(want, &wl);
            for (int q = 1; q < nk; q++)
              if (sm.ord_hash[q] == wh && sm.ord_len[q] == wl && d_streq(sm.names + sm.sh[sm.order[q]].sh_name, want)) return q;
            return -1;
          };
          if (d_streq(t, ".plt")) {
            target = find_name(".got.plt");
            if (target < 0) target = find_name(".got");
          } else {
            target = find_name(t);
          }
        }
        // SHF_INFO_LINK and sh_info exist in the output exactly when BFD finds the target section
        h.sh_flags &= ~(uint64_t)SHF_INFO_LINK;
        h.sh_info = 0;
        if (target >= 0) { h.sh_info = (uint32_t)target; h.sh_flags |= SHF_INFO_LINK; }
      }
      if (h.sh_flags & (0x10 | 0x20)) h.sh_entsize &= 0xffffffffu;  // SHF_MERGE/STRINGS: BFD carries entsize in an unsigned int
      switch (h.sh_type) {  // elf_fake_sections(): entsize of the types BFD knows
        case SHT_INIT_ARRAY: case SHT_FINI_ARRAY: case SHT_PREINIT_ARRAY: case 19 /* RELR */: h.sh_entsize = 8; break;
        case SHT_HASH: h.sh_entsize = 4; break;
        case SHT_DYNAMIC: h.sh_entsize = 16; break;
        case SHT_GNU_HASH: case SHT_GNU_VERDEF: case SHT_GNU_VERNEED: h.sh_entsize = 0; break;
        default: break;
      }
    }
    *reinterpret_cast<Shdr *>(scr + SCR_SHDR + (uint32_t)k * 64) = h;
    // for phase L: where this position's contents go, if it has any
    sm.pos_off[k] = (k >= 1 && k < nk && h.sh_type != SHT_NOBITS && h.sh_size != 0) ? h.sh_offset : ~0ull;
  }
  if (tid >= 96 && tid - 96 < phnum && sm.pkeep[tid - 96]) {   // warp 3: the surviving program headers, compacted
    const int j = tid - 96;
    int slot = 0;
    for (int l = 0; l < j; l++) slot += sm.pkeep[l];
    const uint64_t *s8 = reinterpret_cast<const uint64_t *>(&sm.nph[j]);
    uint64_t *d8 = reinterpret_cast<uint64_t *>(scr + 64 + (uint32_t)slot * 56);
    for (int q = 0; q < 7; q++) d8[q] = s8[q];
  }
  if (tid == 95) {
    Ehdr ne = sm.eh;
    ne.e_shoff = sm.new_shoff;
    ne.e_shnum = (uint16_t)(nk + 1);
    ne.e_shstrndx = (uint16_t)nk;
    ne.e_phnum = (uint16_t)new_phnum;
    *reinterpret_cast<Ehdr *>(scr + SCR_EHDR) = ne;
  }

  LB2_T(7);
  // ---- L. extents: every output byte is produced exactly once -- copied from the input arena,
  //         copied from the scratch slot, or zero-filled (BFD leaves gaps as file holes).
  // pieces = [headers] + content sections sorted by new offset + [.shstrtab] + [section table].  Rank sort,
  // one kept section per thread (the keys are distinct unless sections overlap, which the gap check rejects).
  __syncthreads();   // pos_off[] of phase K
  {
    bool content = false;
    int i = 0;
    if (tid >= 1 && tid < nk) { i = sm.order[tid]; content = sm.pos_off[tid] != ~0ull; }
    if (content) {
      const uint64_t mine = sm.pos_off[tid];
      int r = 0;
#pragma unroll 4
      for (int q = 1; q < nk; q++) {
        const uint64_t o = sm.pos_off[q];   // ~0 for positions without contents: never "before"
        r += (o < mine) || (o == mine && q < tid);
      }
      sm.piece[1 + r] = (uint8_t)i;
    }
    if (tid < 64) {
      const unsigned m = __ballot_sync(0xffffffffu, content);
      if (lane == 0) sm.keep32[warp] = __popc(m);   // (keep32 is free: keepmask lives in registers)
    }
  }
  __syncthreads();
  if (warp == 0) {
    const int np = (int)(sm.keep32[0] + sm.keep32[1]);
    const int total_pieces = np + 3;  // + headers, .shstrtab, section table
    auto piece_of = [&](int q, uint64_t &src, uint64_t &dst, uint64_t &len) {
      if (q == 0) { src = reinterpret_cast<uint64_t>(scr + SCR_EHDR); dst = 0; len = 64 + (uint64_t)new_phnum * 56; }
      else if (q <= np) { const int i = sm.piece[q]; src = sm.src_addr[i]; dst = sm.new_off[i]; len = sm.new_size[i]; }
      else if (q == np + 1) { src = reinterpret_cast<uint64_t>(scr + SCR_STR); dst = sm.shstr_off; len = new_strsz; }
      else { src = reinterpret_cast<uint64_t>(scr + SCR_SHDR); dst = sm.new_shoff; len = (uint64_t)(nk + 1) * 64; }
    };
    ExtWork &x = sm.u.x;
    int ne = 0, bad = 0;
    unsigned long long big0 = 0, big1 = 0, big2 = 0;   // extents with > 64 tiles (bit = extent index, MAX_EXT = 140)
    unsigned long long vb0 = 0, vb1 = 0, vb2 = 0;      // ... of which more than BIG_EXT_TILES: candidates for the expand kernel
    unsigned long long md0 = 0, md1 = 0, md2 = 0;      // extents with 5..64 tiles: one warp each (<= 4 tiles: one thread each)
    auto mark_big = [&](int at, uint32_t c) {
      if (at < 64) big0 |= 1ull << at; else if (at < 128) big1 |= 1ull << (at - 64); else big2 |= 1ull << (at - 128);
      if (c > BIG_EXT_TILES) { if (at < 64) vb0 |= 1ull << at; else if (at < 128) vb1 |= 1ull << (at - 64); else vb2 |= 1ull << (at - 128); }
    };
    auto mark_mid = [&](int at) { if (at < 64) md0 |= 1ull << at; else if (at < 128) md1 |= 1ull << (at - 64); else md2 |= 1ull << (at - 128); };
    unsigned long long copy_bytes = 0;
    uint32_t running = 0;
    for (int q0 = 0; q0 < total_pieces; q0 += 32) {
      const int q = q0 + lane;
      uint64_t src = 0, dst = 0, len = 0, prev_end = 0;
      const bool valid = q < total_pieces;
      if (valid) {
        piece_of(q, src, dst, len);
        if (q > 0) { uint64_t ps, pd, pl; piece_of(q - 1, ps, pd, pl); prev_end = pd + pl; }
        if (dst < prev_end) bad = 1;   // overlapping output ranges: not a layout BFD would write
        if (q == total_pieces - 1 && dst + len != sm.total) bad = 1;  // the section table must end the file
      }
      const bool gap = valid && dst > prev_end;
      const unsigned gm = __ballot_sync(0xffffffffu, gap);
      // tile counts ride along: every extent gets its slot range in the file's tile list by a shuffle scan
      uint32_t cnt_gap = 0, cnt = 0;
      if (gap) cnt_gap = (uint32_t)((dst - 1) / TILE_BYTES - prev_end / TILE_BYTES + 1);
      if (valid && len) cnt = (uint32_t)((dst + len - 1) / TILE_BYTES - dst / TILE_BYTES + 1);
      uint32_t inc = cnt_gap + cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      if (valid) {
        int at = ne + lane + __popc(gm & ((1u << lane) - 1));
        uint32_t t0 = running + inc - (cnt_gap + cnt);
        if (gap) {  // file hole
          x.src[at] = 0; x.dst[at] = prev_end; x.len[at] = dst - prev_end; x.tiles[at] = t0; t0 += cnt_gap;
          if (cnt_gap > 64) mark_big(at, cnt_gap); else if (cnt_gap > 4) mark_mid(at);
          at++;
        }
        x.src[at] = src; x.dst[at] = dst; x.len[at] = len; x.tiles[at] = t0;
        if (cnt > 64) mark_big(at, cnt); else if (cnt > 4) mark_mid(at);
        copy_bytes += len;
      }
      running += __shfl_sync(0xffffffffu, inc, 31);
      const int nvalid = total_pieces - q0 < 32 ? total_pieces - q0 : 32;
      ne += nvalid + __popc(gm);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      copy_bytes += __shfl_xor_sync(0xffffffffu, copy_bytes, o);
      big0 |= __shfl_xor_sync(0xffffffffu, big0, o);
      big1 |= __shfl_xor_sync(0xffffffffu, big1, o);
      big2 |= __shfl_xor_sync(0xffffffffu, big2, o);
      vb0 |= __shfl_xor_sync(0xffffffffu, vb0, o);
      vb1 |= __shfl_xor_sync(0xffffffffu, vb1, o);
      vb2 |= __shfl_xor_sync(0xffffffffu, vb2, o);
      md0 |= __shfl_xor_sync(0xffffffffu, md0, o);
      md1 |= __shfl_xor_sync(0xffffffffu, md1, o);
      md2 |= __shfl_xor_sync(0xffffffffu, md2, o);
    }
    const bool any_bad = __ballot_sync(0xffffffffu, bad) != 0;
    if (lane == 0) {
      sm.n_ext = ne; sm.copy_bytes = copy_bytes; sm.n_tiles = running; sm.big_ext[0] = big0; sm.big_ext[1] = big1; sm.big_ext[2] = big2;
      sm.vbig_ext[0] = vb0; sm.vbig_ext[1] = vb1; sm.vbig_ext[2] = vb2; sm.vbig_taken = 0;
      sm.mid_ext[0] = md0; sm.mid_ext[1] = md1; sm.mid_ext[2] = md2;
      if (any_bad) sm.fail = ST_UNSUPPORTED_LAYOUT;
      else sm.tile_base = atomicAdd(&a.ctr->n_tiles, (unsigned long long)running);   // one atomicAdd per file reserves the range
    }
  }
  __syncthreads();
  if (sm.fail) LB2_REJECT(sm.fail);
  if (a.up_ranges && tid == 0) {
    // host-buffer pipeline: which parts of this file have to cross the bus.  Small files go whole (one DMA);
    // otherwise the input-arena sources of the copy extents, in output order, neighbours closer than 4 KB merged.
    auto push = [&](uint64_t off, uint64_t len) {
      const unsigned int at = atomicAdd(&a.ctr->n_ranges, 1u);
      if (at < a.up_cap) { a.up_ranges[at].off = off; a.up_ranges[at].len = len; }
      else a.ctr->ranges_overflow = 1;
    };
    if (n <= 65536) push(base, n);
    else {
      const uint64_t lo = reinterpret_cast<uint64_t>(in), hi = lo + n;
      uint64_t rs = 0, re = 0;
      for (int e = 0; e < sm.n_ext; e++) {
        const uint64_t s0 = sm.u.x.src[e], l0 = sm.u.x.len[e];
        if (s0 < lo || s0 >= hi || l0 == 0) continue;   // zero fill or regenerated table
        if (re && s0 >= rs && s0 <= re + 4096) { if (s0 + l0 > re) re = s0 + l0; }
        else { if (re) push(base + (rs - lo), re - rs); rs = s0; re = s0 + l0; }
      }
      if (re) push(base + (rs - lo), re - rs);
    }
  }

  LB2_T(8);
  // ---- M. tiles: small extents go one per warp, big ones are written by the whole CTA
  const int n_ext = sm.n_ext;
  const unsigned long long tile_base = sm.tile_base;
  if (tile_base + sm.n_tiles > a.tile_cap) {
    if (tid == 0) { a.ctr->overflow = 1; a.status[f] = ST_PLANNER_LIMIT; a.out_size[f] = 0; }
    return;
  }
  auto emit_tiles = [&](int e, const uint32_t k0, const uint32_t step) {
    const uint64_t d = sm.u.x.dst[e], l = sm.u.x.len[e], s = sm.u.x.src[e];
    if (l == 0) return;
    const uint64_t t0 = d / TILE_BYTES;
    const uint32_t cnt = (uint32_t)((d + l - 1) / TILE_BYTES - t0 + 1);
    Tile *out = a.tiles + tile_base + sm.u.x.tiles[e];
    const uint64_t sbase = s - d;   // source address of output byte b is sbase + b (copies; zero fills keep src = 0)
    if (k0 == 0) {                  // the two edge tiles are the only clipped ones
      Tile t;
      t.file = f;
      const uint64_t e0 = (t0 + 1) * TILE_BYTES < d + l ? (t0 + 1) * TILE_BYTES : d + l;
      t.src = s; t.dst_rel = d; t.len = (uint32_t)(e0 - d);
      out[0] = t;
      if (cnt > 1) {
        const uint64_t b = (t0 + cnt - 1) * TILE_BYTES;
        t.src = s ? sbase + b : 0; t.dst_rel = b; t.len = (uint32_t)(d + l - b);
        out[cnt - 1] = t;
      }
    }
#pragma unroll 4
    for (uint32_t k = 1 + k0; k + 1 < cnt; k += step) {   // interior tiles: whole, aligned
      const uint64_t b = (t0 + k) * TILE_BYTES;
      Tile t;
      t.src = s ? sbase + b : 0;
      t.dst_rel = b;
      t.len = TILE_BYTES;
      t.file = f;
      out[k] = t;
    }
  };
  // extents with more than BIG_EXT_TILES tiles are only recorded: the expand kernel writes their descriptors with the
  // whole grid (if its list is full the CTA writes them itself, as it does the other big ones)
  if (a.big && (sm.vbig_ext[0] | sm.vbig_ext[1] | sm.vbig_ext[2])) {
    if (tid == 0) {
      // one record per <= BIG_PART_TILES tiles, so that a 900 MB section becomes 27 work items for the grid
      unsigned nrec = 0;
      for (int w = 0; w < 3; w++)
        for (unsigned long long m = sm.vbig_ext[w]; m; m &= m - 1) {
          const int e = 64 * w + __ffsll((long long)m) - 1;
          const uint64_t d = sm.u.x.dst[e], l = sm.u.x.len[e];
          const uint32_t cnt = (uint32_t)((d + l - 1) / TILE_BYTES - d / TILE_BYTES + 1);
          nrec += (cnt + BIG_PART_TILES - 1) / BIG_PART_TILES;
        }
      const unsigned at0 = atomicAdd(&a.ctr->n_big, nrec);
      if (at0 + nrec <= a.big_cap) {
        unsigned at = at0;
        for (int w = 0; w < 3; w++)
          for (unsigned long long m = sm.vbig_ext[w]; m; m &= m - 1) {
            const int e = 64 * w + __ffsll((long long)m) - 1;
            const uint64_t sr = sm.u.x.src[e], d = sm.u.x.dst[e], l = sm.u.x.len[e], t0 = d / TILE_BYTES;
            const uint32_t cnt = (uint32_t)((d + l - 1) / TILE_BYTES - t0 + 1);
            for (uint32_t k0 = 0; k0 < cnt; k0 += BIG_PART_TILES) {
              const uint32_t k1 = k0 + BIG_PART_TILES < cnt ? k0 + BIG_PART_TILES : cnt;
              const uint64_t b0 = k0 ? (t0 + k0) * TILE_BYTES : d, b1 = k1 < cnt ? (t0 + k1) * TILE_BYTES : d + l;
              BigExt r;
              r.src = sr ? sr + (b0 - d) : 0; r.dst = b0; r.len = b1 - b0;
              r.tile_index = tile_base + sm.u.x.tiles[e] + k0; r.file = f; r.pad = 0;
              a.big[at++] = r;
            }
          }
        sm.vbig_taken = 1;
      }
    }
    __syncthreads();
  }
  const bool handed_over = sm.vbig_taken != 0;
  // extents of up to 4 tiles (most of them: headers, tables, small sections, gaps): one THREAD each; 5..64 tiles: one
  // warp each; the few with hundreds of tiles were set aside by phase L and are written by the whole CTA
  const unsigned long long big0 = sm.big_ext[0], big1 = sm.big_ext[1], big2 = sm.big_ext[2];
  for (int e = tid; e < n_ext; e += PLAN_THREADS) {
    const uint64_t d = sm.u.x.dst[e], l = sm.u.x.len[e], s = sm.u.x.src[e];
    if (l == 0) continue;
    const uint32_t cnt = (uint32_t)((d + l - 1) / TILE_BYTES - d / TILE_BYTES + 1);
    if (cnt > 4) continue;
    Tile *out = a.tiles + tile_base + sm.u.x.tiles[e];
    for (uint32_t k = 0; k < cnt; k++) out[k] = extent_tile(s, d, l, f, k);
  }
  for (int w = 0; w < 3; w++)
    for (unsigned long long m = sm.mid_ext[w]; m; m &= m - 1) {
      const int e = 64 * w + __ffsll((long long)m) - 1;
      if ((e & (PLAN_THREADS / 32 - 1)) == warp) emit_tiles(e, lane, 32);
    }
  const unsigned long long skip0 = handed_over ? sm.vbig_ext[0] : 0, skip1 = handed_over ? sm.vbig_ext[1] : 0, skip2 = handed_over ? sm.vbig_ext[2] : 0;
  for (unsigned long long m = big0 & ~skip0; m; m &= m - 1) emit_tiles(__ffsll((long long)m) - 1, tid, PLAN_THREADS);
  for (unsigned long long m = big1 & ~skip1; m; m &= m - 1) emit_tiles(64 + __ffsll((long long)m) - 1, tid, PLAN_THREADS);
  for (unsigned long long m = big2 & ~skip2; m; m &= m - 1) emit_tiles(128 + __ffsll((long long)m) - 1, tid, PLAN_THREADS);
  LB2_T(9);
#ifdef LB2_PLAN_TIMING
  if (tid == 0 && f == 0) {
    printf("plan timing (cycles): A=%lld B=%lld C=%lld D+F1=%lld [w0 PH=%lld w1 notes=%lld w2 strtab=%lld] join=%lld I=%lld K=%lld L=%lld M=%lld total=%lld\n",
           t_[1] - t_[0], t_[2] - t_[1], t_[3] - t_[2], t_[4] - t_[3], sm.t_warp[0] - t_[4], sm.t_warp[1] - t_[4], sm.t_warp[2] - t_[4],
           t_[5] - t_[4], t_[6] - t_[5], t_[7] - t_[6], t_[8] - t_[7], t_[9] - t_[8], t_[9] - t_[0]);
  }
#endif
  if (tid == 0) {
    a.out_size[f] = sm.total;
    a.status[f] = ST_OK;
    atomicAdd(&a.ctr->copy_bytes, (unsigned long long)sm.copy_bytes);
    atomicAdd(&a.ctr->out_bytes, (unsigned long long)sm.total);
    atomicAdd(&a.ctr->header_bytes, (unsigned long long)sm.hdr_bytes);
    atomicAdd(&a.ctr->in_bytes, (unsigned long long)n);
    atomicAdd(&a.ctr->n_ok, 1u);
  }
}

#ifndef LB2_HOST_EMULATION  // (the CPU warp emulator under tests/emu compiles only the plan kernel)
// ---------------------------------------------------------------- output offsets
// Block 0: exclusive scan of the 256-byte-rounded output sizes -- where each stripped file starts in the output
// arena (one CTA; n_files is at most a few 10^5).
// Blocks 1..: the tiles of the very big extents the plan kernel only recorded (PlanArgs::big), one record of at most
// BIG_PART_TILES tiles per CTA-iteration: the 54 000 descriptors of a 900 MB section are written by the whole GPU
// while block 0 scans -- same launch, so small batches pay nothing for it.
__global__ void __launch_bounds__(1024) lb2_scan_kernel(const uint64_t *out_size, uint64_t *out_off, uint32_t n,
                                                         uint64_t out_cap, BatchCounters *ctr, const BigExt *big, uint32_t big_cap,
                                                         Tile *tiles) {
  if (blockIdx.x > 0) {
    if (!big || ctr->overflow) return;
    const uint32_t nb = ctr->n_big < big_cap ? ctr->n_big : big_cap;
    for (uint32_t e = blockIdx.x - 1; e < nb; e += gridDim.x - 1) {
      const BigExt r = big[e];
      const uint32_t cnt = (uint32_t)((r.dst + r.len - 1) / TILE_BYTES - r.dst / TILE_BYTES + 1);
      Tile *out = tiles + r.tile_index;
      for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) out[k] = extent_tile(r.src, r.dst, r.len, r.file, k);
    }
    return;
  }
  __shared__ uint64_t warp_excl[32];
  __shared__ uint64_t carry, block_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + tid;
    const uint64_t v = i < n ? ((out_size[i] + 255) & ~255ull) : 0;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_excl[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const uint64_t w = warp_excl[lane];
      uint64_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
      warp_excl[lane] = winc - w;
      if (lane == 31) block_total = winc;
    }
    __syncthreads();
    if (i < n) out_off[i] = carry + warp_excl[wid] + inc - v;
    __syncthreads();
    if (tid == 0) carry += block_total;
    __syncthreads();
  }
  if (tid == 0) {
    out_off[n] = carry;
    if (carry > out_cap) ctr->overflow = 1;
  }
}

void launch_plan(const PlanArgs &a, cudaStream_t s) {
  if (!a.n_files) return;
  lb2_plan_kernel<<<a.n_files, PLAN_THREADS, 0, s>>>(a);  // one CTA per file, one launch for every file
}
void launch_scan(const uint64_t *out_size, uint64_t *out_off, uint32_t n, uint64_t out_cap, BatchCounters *ctr, const BigExt *big,
                 uint32_t big_cap, Tile *tiles, int expand_ctas, cudaStream_t s) {
  lb2_scan_kernel<<<1 + (big ? expand_ctas : 0), 1024, 0, s>>>(out_size, out_off, n, out_cap, ctr, big, big_cap, tiles);
}

#endif  // LB2_HOST_EMULATION

}  // namespace lb2
