/*
 * strip_oracle.c -- CPU restatement of what GNU `strip` (Binutils/BFD 2.42) does to an
 * ELF64 little-endian shared object / executable.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may build, load or run it.
 * The product (lambdipy_b200/csrc/ *.cu) is an independent CUDA implementation; it never calls
 * into this file and fails loudly when its CUDA library is missing.
 *
 * What is restated and why it is not "the reference's source":
 *   The reference (customink/lambdipy) performs this step by shelling out to the external
 *   binary `strip` -- /root/reference/lambdipy/project_build.py:260
 *       find {install_dir}/ -name "*.so" | xargs strip
 *   The byte-level algorithm therefore lives in a third-party dependency that is NOT in
 *   /root/reference and is NOT version-pinned by it (no lockfile; whatever `strip` is on PATH,
 *   project_build.py:191-192,268).  BASELINE.json fixes the parity target to this image's
 *   /usr/bin/strip = GNU Binutils 2.42.  Binutils source is not available offline, so the
 *   rules below restate BFD's published behaviour (objcopy.c: is_strip_section,
 *   merge_gnu_build_notes; bfd/elf.c: assign_section_numbers,
 *   assign_file_positions_for_load_sections / _non_load_sections, copy_elf_program_header;
 *   bfd/elf-strtab.c: _bfd_elf_strtab_finalize) and are PINNED by differential testing against
 *   the real binary (tests/test_oracle_vs_gnu_strip.py, tests/golden/).  Rule numbers R1..R12
 *   refer to /root/repo/SURVEY.md section 8(c).
 *
 * Build: see oracle/Makefile (gcc -O2 -shared).  API at the bottom of this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ---------------------------------------------------------------- ELF64-LE structures */
typedef struct {
  uint8_t e_ident[16];
  uint16_t e_type, e_machine;
  uint32_t e_version;
  uint64_t e_entry, e_phoff, e_shoff;
  uint32_t e_flags;
  uint16_t e_ehsize, e_phentsize, e_phnum, e_shentsize, e_shnum, e_shstrndx;
} Ehdr;
typedef struct {
  uint32_t p_type, p_flags;
  uint64_t p_offset, p_vaddr, p_paddr, p_filesz, p_memsz, p_align;
} Phdr;
typedef struct {
  uint32_t sh_name, sh_type;
  uint64_t sh_flags, sh_addr, sh_offset, sh_size;
  uint32_t sh_link, sh_info;
  uint64_t sh_addralign, sh_entsize;
} Shdr;

enum { SHT_NULL = 0, SHT_PROGBITS = 1, SHT_SYMTAB = 2, SHT_STRTAB = 3, SHT_RELA = 4, SHT_NOTE = 7,
       SHT_NOBITS = 8, SHT_REL = 9, SHT_DYNSYM = 11, SHT_GROUP = 17, SHT_SYMTAB_SHNDX = 18 };
enum { SHF_WRITE = 1, SHF_ALLOC = 2, SHF_INFO_LINK = 0x40, SHF_TLS = 0x400 };
enum { PT_NULL = 0, PT_LOAD = 1, PT_DYNAMIC = 2, PT_INTERP = 3, PT_NOTE = 4, PT_PHDR = 6, PT_TLS = 7,
       PT_GNU_EH_FRAME = 0x6474e550, PT_GNU_STACK = 0x6474e551, PT_GNU_RELRO = 0x6474e552,
       PT_GNU_PROPERTY = 0x6474e553 };
enum { ET_EXEC = 2, ET_DYN = 3 };

/* Return codes.  0 = stripped.  Positive = a file class this restatement does not cover (the
 * product routes the same classes to its host-`strip` fallback).  Negative = malformed. */
enum {
  LBO_OK = 0,
  LBO_NOT_ELF = 1,          /* GNU strip: "file format not recognized" */
  LBO_NOT_ELF64LE = 2,
  LBO_BAD_TYPE = 3,         /* ET_REL / ET_CORE / ...: different BFD code path */
  LBO_NO_SECTIONS = 4,      /* GNU strip: "has no sections" */
  LBO_XINDEX = 5,
  LBO_UNSUPPORTED_LAYOUT = 6,
  LBO_BAD_NOTES = 7,        /* build-attribute notes objcopy would refuse to merge in a way we don't model */
  LBO_MALFORMED = -1,
  LBO_NOMEM = -2
};

#define LBO_FLAG_NO_MERGE_NOTES 1u

static int lbo_debug_line = 0; /* last UNSUP() site, for diagnostics (lbo_last_unsupported_line) */
#define UNSUP() do { rc = LBO_UNSUPPORTED_LAYOUT; lbo_debug_line = __LINE__; goto done; } while (0)

static uint64_t align_up(uint64_t v, uint64_t a) { return a > 1 ? (v + a - 1) / a * a : v; }

static int has_prefix(const char *s, const char *p) { return strncmp(s, p, strlen(p)) == 0; }

/* ---------------------------------------------------------------- R1: keep / drop
 * objcopy.c is_strip_section() under strip_all/strip_unneeded, plus BFD's own handling of the
 * symbol/string tables: .symtab, non-alloc STRTABs (.strtab, old .shstrtab) are regenerated or
 * omitted by BFD; sections BFD flags SEC_DEBUGGING (elf.c _bfd_elf_make_section_from_shdr: the
 * name prefixes below, only when not SHF_ALLOC) are removed. */
static int is_debug_name(const char *n) {
  return has_prefix(n, ".debug") || has_prefix(n, ".zdebug") || has_prefix(n, ".gnu.debuglto_.debug_") ||
         has_prefix(n, ".gnu.linkonce.wi.") || has_prefix(n, ".line") || has_prefix(n, ".stab") ||
         strcmp(n, ".gdb_index") == 0;
}

/* ---------------------------------------------------------------- R6: .shstrtab (elf-strtab.c) */
typedef struct {
  const char *s;
  int len;      /* strlen, without NUL */
  int order;    /* insertion order */
  int host;     /* index of entry this one is a suffix of, or -1 */
  uint32_t off; /* final offset */
} StrEnt;

static int strrev_cmp(const void *pa, const void *pb) {
  const StrEnt *a = *(const StrEnt *const *)pa, *b = *(const StrEnt *const *)pb;
  int l = a->len < b->len ? a->len : b->len;
  const unsigned char *s = (const unsigned char *)a->s + a->len - 1;
  const unsigned char *t = (const unsigned char *)b->s + b->len - 1;
  while (l--) {
    if (*s != *t) return (int)*s - (int)*t;
    s--, t--;
  }
  return a->len - b->len;
}

/* Build the tail-merged table for `n` unique strings (in insertion order).  Returns the blob. */
static uint8_t *build_strtab(StrEnt *e, int n, uint32_t *size_out) {
  StrEnt **arr = (StrEnt **)malloc(sizeof(StrEnt *) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) { arr[i] = &e[i]; e[i].host = -1; }
  if (n) {
    qsort(arr, (size_t)n, sizeof(StrEnt *), strrev_cmp); /* names are unique => total order */
    StrEnt *cur = arr[n - 1];
    for (int i = n - 2; i >= 0; i--) {
      StrEnt *c = arr[i];
      /* is_suffix: cur is strictly longer and ends with c */
      if (cur->len > c->len && memcmp(cur->s + (cur->len - c->len), c->s, (size_t)c->len) == 0)
        c->host = (int)(cur - e);
      else
        cur = c;
    }
  }
  uint32_t size = 1;
  for (int i = 0; i < n; i++)
    if (e[i].host < 0) { e[i].off = size; size += (uint32_t)e[i].len + 1; }
  for (int i = 0; i < n; i++)
    if (e[i].host >= 0) { StrEnt *h = &e[e[i].host]; e[i].off = h->off + (uint32_t)(h->len - e[i].len); }
  uint8_t *blob = (uint8_t *)calloc(size, 1);
  for (int i = 0; i < n; i++)
    if (e[i].host < 0) memcpy(blob + e[i].off, e[i].s, (size_t)e[i].len);
  free(arr);
  *size_out = size;
  return blob;
}

/* ---------------------------------------------------------------- R9: objcopy merge_gnu_build_notes */
typedef struct {
  uint32_t namesz, descsz, type, padded_namesz;
  const uint8_t *name;
  uint64_t start, end;
  uint32_t seq; /* position in the input, for diagnostics */
} BNote;

#define NT_OPEN 0x100u
#define NT_FUNC 0x101u

static int note_is_open(const BNote *n) { return n->type == NT_OPEN; }
static int note_is_deleted(const BNote *n) { return n->type == 0; }
static int note_is_version(const BNote *n) {
  return n->namesz > 4 && n->name[0] == 'G' && n->name[1] == 'A' && n->name[2] == '$' && n->name[3] == 1;
}

static int cmp_notes_by_attr(const void *a, const void *b) {
  const BNote *p1 = (const BNote *)a, *p2 = (const BNote *)b;
  uint32_t l = (p1->namesz < p2->namesz ? p1->namesz : p2->namesz) - 3;
  int c = memcmp(p1->name + 3, p2->name + 3, l);
  if (c) return c;
  if (p1->end < p2->start) return -1;
  if (p1->start > p2->end) return 1;
  if (p1->start < p2->start) return -1;
  if (p1->end > p2->end) return 1;
  if (p1->end < p2->end) return -1;
  if (note_is_open(p1) && !note_is_open(p2)) return -1;
  if (!note_is_open(p1) && note_is_open(p2)) return 1;
  return 0;
}

static int cmp_notes_by_addr(const void *a, const void *b) {
  const BNote *p1 = (const BNote *)a, *p2 = (const BNote *)b;
  /* OPEN notes first, then by start address, larger ranges first (probe: random OPEN/FUNC mixes) */
  if (note_is_open(p1) && !note_is_open(p2)) return -1;
  if (!note_is_open(p1) && note_is_open(p2)) return 1;
  if (p1->start < p2->start) return -1;
  if (p1->start > p2->start) return 1;
  if (p1->end > p2->end) return -1;
  if (p1->end < p2->end) return 1;
  /* ties keep the order of the first sort (the merge sort is stable) -- probe: a corrupted "HA$<version>"
   * note stays between two version notes of the same range */
  return 0;
}

/* objcopy sorts with libc qsort().  compare_gnu_build_notes is not antisymmetric for nested
 * ranges ([0x1000,0x1400] and [0x1200,0x1300] each compare "less" than the other), so the outcome
 * depends on the comparison sequence of the C library.  glibc 2.39 (this image; its qsort was
 * probed stable for n up to 1.2e6) uses the classic top-down merge sort of stdlib/msort.c:
 * n1 = n / 2, sort both halves, merge taking the left element while cmp(left, right) <= 0.
 * Restated here so the oracle does not depend on which libc it is linked against. */
static void msort_notes(BNote *b, size_t n, BNote *tmp, int (*cmp)(const void *, const void *)) {
  if (n <= 1) return;
  size_t n1 = n / 2, n2 = n - n1;
  BNote *b1 = b, *b2 = b + n1, *t = tmp;
  msort_notes(b1, n1, tmp, cmp);
  msort_notes(b2, n2, tmp, cmp);
  while (n1 > 0 && n2 > 0) {
    if (cmp(b1, b2) <= 0) { *t++ = *b1++; n1--; }
    else { *t++ = *b2++; n2--; }
  }
  if (n1 > 0) memcpy(t, b1, n1 * sizeof(BNote));
  memcpy(b, tmp, (n - n2) * sizeof(BNote));
}

/* Returns new size (<= size) written into `out` (capacity 2*size); or `size` with out == copy. */
static uint64_t merge_build_notes(const uint8_t *in, uint64_t size, uint8_t *out, int *err) {
  *err = 0;
  memcpy(out, in, size);
  if (size == 0) return size;   /* objcopy skips empty note sections; 1..11 bytes are "excess data" -> corrupt */
  BNote *notes = (BNote *)calloc(size / 12 + 1, sizeof(BNote));
  BNote *pn = notes;
  uint64_t remain = size;
  const uint8_t *p = in;
  unsigned v1 = 0, v2 = 0, v3 = 0;
  uint64_t prev_func_start = 0, prev_open_start = 0, prev_func_end = 0, prev_open_end = 0;
  while (remain >= 12) {
    uint64_t start, end;
    memcpy(&pn->namesz, p, 4);
    memcpy(&pn->descsz, p + 4, 4);
    memcpy(&pn->type, p + 8, 4);
    pn->padded_namesz = (pn->namesz + 3) & ~3u;
    if (pn->namesz > 0xfffffff0u) goto bad;   /* objcopy computes in unsigned long: "note too big" */
    if (((pn->descsz + 3) & ~3u) != pn->descsz) goto bad;
    if (pn->type != NT_OPEN && pn->type != NT_FUNC) goto bad;
    if ((uint64_t)pn->padded_namesz + pn->descsz + 12 > remain) goto bad;
    if (pn->namesz < 3) goto bad; /* objcopy accepts 2 and then memcmp()s namesz - 3 bytes: out of contract */
    pn->name = p + 12;
    const uint8_t *desc = p + 12 + pn->padded_namesz;
    remain -= 12 + (uint64_t)pn->padded_namesz + pn->descsz;
    p += 12 + (uint64_t)pn->padded_namesz + pn->descsz;
    if (pn->namesz > 2 && pn->name[0] == '$' && pn->name[1] == 1 && pn->name[2] == '1') v1++;
    else if (note_is_version(pn)) {
      if (pn->name[4] == '2') v2++;
      else if (pn->name[4] == '3') v3++;
      else goto bad;
    }
    switch (pn->descsz) {
      case 0: start = end = 0; break;
      case 4: { uint32_t a; memcpy(&a, desc, 4); start = a; end = (uint64_t)-1; break; }
      case 8: { uint32_t a, b; memcpy(&a, desc, 4); memcpy(&b, desc + 4, 4); start = a; end = b; break; }
      case 16: memcpy(&start, desc, 8); memcpy(&end, desc + 8, 8); break;
      default: goto bad;
    }
    if (start > end) start = end;
    if (note_is_open(pn)) {
      if (start) prev_open_start = start;
      pn->start = prev_open_start;
      if (end) prev_open_end = end;
      pn->end = prev_open_end;
    } else {
      if (start) prev_func_start = start;
      pn->start = prev_func_start;
      if (end) prev_func_end = end;
      pn->end = prev_func_end;
    }
    if (pn->name[pn->namesz - 1] != 0) goto bad;
    pn->seq = (uint32_t)(pn - notes);
    pn++;
  }
  BNote *pend = pn;
  if (remain != 0) goto bad;
  if (v1 == 0 && v2 == 0 && v3 == 0) v3 = 2; /* "version note missing - assuming version 3" */
  if ((v1 && v2) || (v1 && v3) || (v2 && v3)) goto bad;
  if (v3 == 0) { free(notes); return size; } /* only v3 notes are merged */

  BNote *tmp = (BNote *)calloc((size_t)(pend - notes) + 1, sizeof(BNote));
  msort_notes(notes, (size_t)(pend - notes), tmp, cmp_notes_by_attr);

  for (pn = notes; pn < pend; pn++) {
    if (note_is_deleted(pn)) continue;
    if (pn->start == pn->end) { pn->type = 0; continue; } /* rule 1: empty range */
    int iter = 0;
    for (BNote *back = pn - 1; back >= notes; back--) {
      if (note_is_deleted(back)) continue;
      if (back->namesz != pn->namesz || memcmp(back->name, pn->name, pn->namesz) != 0) break;
      if (back->start == pn->start && back->end == pn->end) { pn->type = 0; break; } /* rule 2 */
      /* a note whose range lies inside an earlier same-name note's range is redundant
       * (objcopy.c contained_by(pnote, back)) */
      if (pn->start >= back->start && pn->end <= back->end) { pn->type = 0; break; }
      /* rule 3: objcopy.c overlaps_or_adjoins(back, pnote), restated as published -- including its
       * inverted gap test: when back ends before pnote starts the ranges are merged iff there IS
       * a gap after rounding back's end up to 16; otherwise they merge unless both ends are equal. */
      {
        int merge;
        if (back->end < pn->start) merge = (((back->end + 15) & ~(uint64_t)15) < pn->start);
        else merge = (back->end != pn->end);
        /* only notes of the same kind are combined (probe: an OPEN and a FUNC note of the same
         * attribute with overlapping ranges both survive; identical / contained ones do not) */
        if (back->type != pn->type) merge = 0;
        if (merge) {
          back->start = back->start < pn->start ? back->start : pn->start;
          back->end = back->end > pn->end ? back->end : pn->end;
          pn->type = 0;
          break;
        }
      }
      if (iter++ > 16) break;
    }
  }

  msort_notes(notes, (size_t)(pend - notes), tmp, cmp_notes_by_addr);
  free(tmp);

  uint8_t *w = out;
  uint64_t prev_start = 0, prev_end = 0;
  for (pn = notes; pn < pend; pn++) {
    if (note_is_deleted(pn)) continue;
    int elide = (pn->start == prev_start && pn->end == prev_end);
    uint32_t dsz = elide ? 0 : 16;
    memcpy(w, &pn->namesz, 4);
    memcpy(w + 4, &dsz, 4);
    memcpy(w + 8, &pn->type, 4);
    w += 12;
    memcpy(w, pn->name, pn->namesz);
    if (pn->namesz < pn->padded_namesz) memset(w + pn->namesz, 0, pn->padded_namesz - pn->namesz);
    w += pn->padded_namesz;
    if (!elide) {
      memcpy(w, &pn->start, 8);
      memcpy(w + 8, &pn->end, 8);
      w += 16;
      prev_start = pn->start;
      prev_end = pn->end;
    }
  }
  uint64_t new_size = (uint64_t)(w - out);
  free(notes);
  if (new_size < size) return new_size;
  memcpy(out, in, size);
  return size;
bad:
  /* objcopy reports the corrupt note and leaves the section as it was (and exits non-zero). */
  free(notes);
  *err = 1;
  memcpy(out, in, size);
  return size;
}

/* ---------------------------------------------------------------- the strip itself */
typedef struct {
  Shdr h;            /* input header */
  const char *name;
  int keep;
  int new_index;
  int seg;           /* index of the PT_LOAD that holds it (alloc sections), or -1 */
  uint64_t new_off;
  uint64_t new_size;
  const uint8_t *data; /* source bytes (input or regenerated) */
  uint8_t *owned;      /* regenerated contents to free */
} Sec;

/* BFD's ELF_SECTION_IN_SEGMENT(sec_hdr, segment) (include/elf/internal.h; check_vma = 1,
 * strict = 0), the predicate copy_elf_program_header() uses to decide which sections a program
 * header carries. */
static uint64_t sec_size_in_seg(const Shdr *s, const Phdr *p) {
  /* .tbss occupies no space except in PT_TLS */
  if ((s->sh_flags & SHF_TLS) && s->sh_type == SHT_NOBITS && p->p_type != PT_TLS) return 0;
  return s->sh_size;
}
static int sec_in_seg(const Shdr *s, const Phdr *p) {
  uint32_t t = p->p_type;
  int tls = (s->sh_flags & SHF_TLS) != 0, alloc = (s->sh_flags & SHF_ALLOC) != 0;
  uint64_t sz = sec_size_in_seg(s, p);
  if (tls) { if (!(t == PT_TLS || t == PT_GNU_RELRO || t == PT_LOAD)) return 0; }
  else if (t == PT_TLS || t == PT_PHDR) return 0;
  if (!alloc && (t == PT_LOAD || t == PT_DYNAMIC || t == PT_GNU_EH_FRAME || t == PT_GNU_STACK || t == PT_GNU_RELRO ||
                 t == 0x6474e554u /* PT_GNU_SFRAME */ || (t >= 0x6474e555u && t <= 0x6474f554u) /* PT_GNU_MBIND */))
    return 0;
  if (s->sh_type != SHT_NOBITS) {
    if (s->sh_offset < p->p_offset) return 0;
    if (s->sh_offset - p->p_offset + sz > p->p_filesz) return 0;
  }
  if (alloc) {
    if (s->sh_addr < p->p_vaddr) return 0;
    if (s->sh_addr - p->p_vaddr + sz > p->p_memsz) return 0;
  }
  if ((t == PT_DYNAMIC || t == PT_NOTE) && s->sh_size == 0 && p->p_memsz != 0) {
    /* no zero-size sections at the start or end of PT_DYNAMIC / PT_NOTE */
    int ok_off = s->sh_type == SHT_NOBITS ||
                 (s->sh_offset > p->p_offset && s->sh_offset - p->p_offset < p->p_filesz);
    int ok_vma = !alloc || (s->sh_addr > p->p_vaddr && s->sh_addr - p->p_vaddr < p->p_memsz);
    if (!(ok_off && ok_vma)) return 0;
  }
  return 1;
}

/* ---------------------------------------------------------------- conservative input gate
 * BFD normalises a number of header fields from its own tables (section type and flags by NAME,
 * sh_link by name lookup of .dynstr/.dynsym, LMA from p_paddr, ...).  On files written by ld, gold,
 * lld, patchelf or objcopy those fields already hold BFD's values, so the rules above never see the
 * difference.  Anything that deviates is declared out of contract (LBO_UNSUPPORTED_LAYOUT -> the
 * product hands the file to the host strip) instead of being guessed at.  Found with
 * oracle/fuzz_vs_gnu.py. */
static int name_is(const char *n, const char *base) {   /* "base" or "base.*" (BFD prefix entries with -2) */
  size_t l = strlen(base);
  return strncmp(n, base, l) == 0 && (n[l] == 0 || n[l] == '.');
}
static int expected_type_by_name(const char *n) {  /* bfd/elf.c special_sections_*; -1: not a special name */
  if (name_is(n, ".bss") || name_is(n, ".tbss") || name_is(n, ".sbss") || name_is(n, ".lbss") || has_prefix(n, ".gnu.linkonce.b") || name_is(n, ".noinit")) return SHT_NOBITS;
  if (strcmp(n, ".comment") == 0 || name_is(n, ".data") || name_is(n, ".data1") || has_prefix(n, ".debug") || strcmp(n, ".fini") == 0 ||
      strcmp(n, ".got") == 0 || strcmp(n, ".init") == 0 || strcmp(n, ".interp") == 0 || has_prefix(n, ".line") || strcmp(n, ".plt") == 0 ||
      name_is(n, ".rodata") || name_is(n, ".rodata1") || name_is(n, ".tdata") || name_is(n, ".text") || name_is(n, ".sdata") ||
      name_is(n, ".ldata") || name_is(n, ".lrodata") || name_is(n, ".persistent") || has_prefix(n, ".gnu.linkonce.wi."))
    return SHT_PROGBITS;
  if (strcmp(n, ".dynamic") == 0) return 6;
  if (strcmp(n, ".dynstr") == 0 || strcmp(n, ".strtab") == 0 || strcmp(n, ".shstrtab") == 0) return SHT_STRTAB;
  if (strcmp(n, ".dynsym") == 0) return SHT_DYNSYM;
  if (strcmp(n, ".symtab") == 0) return SHT_SYMTAB;
  if (name_is(n, ".fini_array")) return 15;
  if (name_is(n, ".init_array")) return 14;
  if (name_is(n, ".preinit_array")) return 16;
  if (strcmp(n, ".gnu.version") == 0) return 0x6fffffff;
  if (strcmp(n, ".gnu.version_d") == 0) return 0x6ffffffd;
  if (strcmp(n, ".gnu.version_r") == 0) return 0x6ffffffe;
  if (strcmp(n, ".gnu.hash") == 0) return 0x6ffffff6;
  if (strcmp(n, ".hash") == 0) return 5;
  if (has_prefix(n, ".note")) return SHT_NOTE;
  if (has_prefix(n, ".rela")) return SHT_RELA;
  if (name_is(n, ".rel")) return SHT_REL;   /* on RELA targets BFD matches ".rel" only as "rel" or "rel.*" (not .relro_padding, .relr.dyn) */
  return -1;
}
static int type_is_known(uint32_t t) {
  switch (t) {
    case SHT_PROGBITS: case SHT_SYMTAB: case SHT_STRTAB: case SHT_RELA: case 5: case 6: case SHT_NOTE: case SHT_NOBITS: case SHT_DYNSYM:
    case 14: case 15: case 16: case 19 /* SHT_RELR */: case 0x6ffffff6: case 0x6ffffffd: case 0x6ffffffe: case 0x6fffffff: case 0x70000001: return 1;
    default: return 0;
  }
}

int lbo_strip(const uint8_t *in, uint64_t n, uint8_t **out_p, uint64_t *out_n, unsigned flags) {
  *out_p = NULL;
  *out_n = 0;
  if (n < 64 || memcmp(in, "\177ELF", 4) != 0) return LBO_NOT_ELF;
  if (in[4] != 2 || in[5] != 1) return LBO_NOT_ELF64LE;
  Ehdr eh;
  memcpy(&eh, in, 64);
  if (eh.e_type != ET_DYN && eh.e_type != ET_EXEC) return LBO_BAD_TYPE;
  if (eh.e_machine != 62 && eh.e_machine != 183) return LBO_UNSUPPORTED_LAYOUT; /* x86-64, aarch64 only */
  /* gate: BFD writes EV_CURRENT / sizeof(Ehdr) itself; a file that says otherwise is not linker output */
  if (eh.e_version != 1 || in[6] != 1 || eh.e_ehsize != 64) return LBO_UNSUPPORTED_LAYOUT;
  if (eh.e_shoff == 0 || eh.e_shnum == 0) return LBO_NO_SECTIONS;
  if (eh.e_shentsize != 64 || (eh.e_phnum && eh.e_phentsize != 56)) return LBO_MALFORMED;
  if (eh.e_shstrndx == 0xffff || eh.e_shnum >= 0xff00 || eh.e_phnum == 0xffff) return LBO_XINDEX;
  uint64_t shnum = eh.e_shnum, phnum = eh.e_phnum;
  if (eh.e_shoff > n || shnum * 64 > n - eh.e_shoff) return LBO_MALFORMED;
  if (eh.e_phoff > n || phnum * 56 > n - eh.e_phoff) return LBO_MALFORMED;
  if (eh.e_shstrndx >= shnum) return LBO_MALFORMED;
  if (phnum && eh.e_phoff != 64) return LBO_UNSUPPORTED_LAYOUT;

  int rc = LBO_OK;
  Sec *S = (Sec *)calloc(shnum, sizeof(Sec));
  Shdr *in_hdr = (Shdr *)calloc(shnum, sizeof(Shdr));
  Phdr *P = (Phdr *)calloc(phnum ? phnum : 1, sizeof(Phdr));
  Phdr *NP = (Phdr *)calloc(phnum ? phnum : 1, sizeof(Phdr));
  int *pkeep = (int *)calloc(phnum ? phnum : 1, sizeof(int));
  int64_t *pshift = (int64_t *)calloc(phnum ? phnum : 1, sizeof(int64_t));
  StrEnt *ents = (StrEnt *)calloc(shnum + 1, sizeof(StrEnt));
  int *order = (int *)calloc(shnum + 1, sizeof(int));
  uint8_t *out = NULL, *strblob = NULL;
  if (!S || !in_hdr || !P || !NP || !pkeep || !pshift || !ents || !order) { rc = LBO_NOMEM; goto done; }

  for (uint64_t i = 0; i < shnum; i++) { memcpy(&S[i].h, in + eh.e_shoff + i * 64, 64); in_hdr[i] = S[i].h; }
  for (uint64_t j = 0; j < phnum; j++) memcpy(&P[j], in + eh.e_phoff + j * 56, 56);
  const Shdr *strh = &S[eh.e_shstrndx].h;
  if (strh->sh_type != SHT_STRTAB || strh->sh_offset > n || strh->sh_size > n - strh->sh_offset || strh->sh_size == 0 ||
      in[strh->sh_offset + strh->sh_size - 1] != 0) { rc = LBO_MALFORMED; goto done; }
  for (uint64_t i = 0; i < shnum; i++) {
    if (S[i].h.sh_name >= strh->sh_size) { rc = LBO_MALFORMED; goto done; }
    S[i].name = (const char *)in + strh->sh_offset + S[i].h.sh_name;
    if (S[i].h.sh_type != SHT_NOBITS && S[i].h.sh_type != SHT_NULL &&
        (S[i].h.sh_offset > n || S[i].h.sh_size > n - S[i].h.sh_offset)) { rc = LBO_MALFORMED; goto done; }
  }

  for (uint64_t j = 0; j < phnum; j++) { /* gate: section LMAs come from p_paddr; Linux objects have paddr == vaddr */
    if (P[j].p_paddr != P[j].p_vaddr) UNSUP();
    /* BFD refuses "a program header with invalid alignment"; loadable segments must be congruent */
    if (P[j].p_align & (P[j].p_align - 1)) UNSUP();
    if (P[j].p_type == PT_LOAD && P[j].p_align > 1 && ((P[j].p_vaddr - P[j].p_offset) & (P[j].p_align - 1))) UNSUP();
    if (P[j].p_type == PT_PHDR && (P[j].p_offset != 64 || P[j].p_filesz != phnum * 56 || P[j].p_memsz != phnum * 56)) UNSUP();
    if (P[j].p_type == PT_GNU_STACK && (P[j].p_offset || P[j].p_vaddr || P[j].p_filesz || P[j].p_memsz)) UNSUP();
  }
  {
    /* gate: loadable segments are ascending and do not overlap, in memory or in the file */
    uint64_t vend = 0, fend = 0;
    int seen = 0;
    for (uint64_t j = 0; j < phnum; j++) {
      if (P[j].p_type != PT_LOAD) continue;
      if (P[j].p_filesz > P[j].p_memsz) UNSUP();
      if (P[j].p_vaddr + P[j].p_memsz < P[j].p_vaddr || P[j].p_offset + P[j].p_filesz < P[j].p_offset) UNSUP();
      if (seen && (P[j].p_vaddr < vend || (P[j].p_filesz && P[j].p_offset < fend))) UNSUP();
      vend = P[j].p_vaddr + P[j].p_memsz;
      if (P[j].p_filesz) fend = P[j].p_offset + P[j].p_filesz;
      seen = 1;
    }
  }
  {
    static const uint8_t zero64[64] = {0};
    if (memcmp(in + eh.e_shoff, zero64, 64) != 0) UNSUP();   /* section 0 must be the all-zero NULL header */
  }

  /* R1 */
  for (uint64_t i = 1; i < shnum; i++) {
    const Shdr *h = &S[i].h;
    int alloc = (h->sh_flags & SHF_ALLOC) != 0;
    int drop = 0;
    if (h->sh_type == SHT_SYMTAB || h->sh_type == SHT_SYMTAB_SHNDX) drop = 1;
    else if (h->sh_type == SHT_STRTAB && !alloc) drop = 1;
    else if (!alloc && is_debug_name(S[i].name)) drop = 1;
    if (h->sh_type == SHT_NULL || h->sh_type == SHT_GROUP) UNSUP();
    /* BFD refuses ("file format not recognized") symbol/reloc/versym tables with a foreign entsize */
    if ((h->sh_type == SHT_DYNSYM || h->sh_type == SHT_SYMTAB || h->sh_type == SHT_RELA) && h->sh_entsize != 24) UNSUP();
    if (h->sh_type == 0x6fffffff && h->sh_entsize != 2) UNSUP();
    if (h->sh_type == 19 && h->sh_entsize != 8) UNSUP();   /* SHT_RELR */
    if (h->sh_type == SHT_REL && h->sh_entsize != 16) UNSUP();
    if (!alloc && (h->sh_type == SHT_REL || h->sh_type == SHT_RELA)) UNSUP();
    /* ---- gate (see above) */
    {
      const uint64_t ALLOWED = 0x1 | 0x2 | 0x4 | 0x10 | 0x20 | 0x40 | 0x400 | 0x800 | 0x200000 | 0x10000000;
      const char *nm = S[i].name;
      int want = expected_type_by_name(nm);
      if (h->sh_flags & ~ALLOWED) UNSUP();
      if (!type_is_known(h->sh_type)) UNSUP();
      if (h->sh_type == SHT_NOBITS && !alloc) UNSUP();
      if (want >= 0 && (uint32_t)want != h->sh_type && !(h->sh_type == 0x70000001 && want == SHT_PROGBITS)) UNSUP();
      if ((h->sh_flags & SHF_INFO_LINK) && h->sh_type != SHT_RELA && h->sh_type != SHT_REL) UNSUP();
      if (h->sh_link >= shnum) UNSUP();
      switch (h->sh_type) {
        case SHT_DYNSYM: case 6: case 0x6ffffffd: case 0x6ffffffe:   /* BFD: sh_link := index of ".dynstr" */
          if (h->sh_link == 0 || strcmp(S[h->sh_link].name, ".dynstr") != 0) UNSUP();
          if (h->sh_type == 6 && h->sh_info != 0) UNSUP();
          if (h->sh_type == SHT_DYNSYM && (h->sh_size % 24 != 0 || h->sh_info > h->sh_size / 24)) UNSUP(); /* BFD refuses */
          break;
        case 5: case 0x6ffffff6: case 0x6fffffff:                     /* BFD: sh_link := index of ".dynsym" */
          if (h->sh_link == 0 || strcmp(S[h->sh_link].name, ".dynsym") != 0 || h->sh_info != 0) UNSUP();
          break;
        case SHT_RELA: case SHT_REL:
          if (h->sh_link != 0 && strcmp(S[h->sh_link].name, ".dynsym") != 0) UNSUP();
          if (h->sh_info >= shnum) UNSUP();
          break;
        case SHT_SYMTAB:                                              /* dropped, but BFD reads it first and refuses a broken one */
          if (h->sh_link == 0 || S[h->sh_link].h.sh_type != SHT_STRTAB || (S[h->sh_link].h.sh_flags & SHF_ALLOC)) UNSUP();
          if (h->sh_size % 24 != 0 || h->sh_info > h->sh_size / 24) UNSUP();
          break;
        case SHT_STRTAB:
          if (h->sh_link != 0 || h->sh_info != 0) UNSUP();
          break;
        default:                                                      /* ordinary sections: BFD writes 0/0 */
          if (h->sh_link != 0 || h->sh_info != 0) UNSUP();
          break;
      }
    }
    /* gate: a TLS NOBITS section (.tbss) must sit where the linker puts it -- at the aligned end of the TLS section
     * before it (or at the start of PT_TLS).  BFD derives the file offset it writes for .tbss from that address
     * (probe: .tbss sh_addr moved by 4 / 16 / 256 bytes -> sh_offset 0x2dc8 / 0x2dd0 / 0x2f90 instead of 0x2db4);
     * only the natural placement is reproduced. */
    if ((h->sh_flags & SHF_TLS) && (h->sh_flags & SHF_ALLOC) && h->sh_type == SHT_NOBITS) {
      uint64_t al = h->sh_addralign ? h->sh_addralign : 1;
      if (al & (al - 1)) UNSUP();
      int64_t pj = -1;
      for (int64_t j = (int64_t)i - 1; j >= 1; j--)
        if ((S[j].h.sh_flags & SHF_TLS) && (S[j].h.sh_flags & SHF_ALLOC)) { pj = j; break; }
      if (pj >= 0) {
        uint64_t pe = S[pj].h.sh_addr + S[pj].h.sh_size;
        if (h->sh_addr < pe || h->sh_addr - pe >= al || (h->sh_addr & (al - 1))) UNSUP();
      } else {
        for (uint64_t j = 0; j < phnum; j++)
          if (P[j].p_type == PT_TLS && h->sh_addr != P[j].p_vaddr) UNSUP();
      }
    }
    S[i].keep = !drop;
    /* BFD keeps alignment as a power of two that the section address honours
     * (probe: doctored sh_addralign 0/3/24/4096 -> min(lowbit(align), lowbit(addr)), 0 -> 1). */
    {
      uint64_t al = h->sh_addralign ? (h->sh_addralign & (~h->sh_addralign + 1)) : 1;
      if (h->sh_addr) { uint64_t lb = h->sh_addr & (~h->sh_addr + 1); if (lb < al) al = lb; }
      S[i].h.sh_addralign = al;
    }
    S[i].seg = -1;
    S[i].data = in + h->sh_offset;
    S[i].new_size = h->sh_size;
  }
  S[0].keep = 1;

  /* R2: output order = input order, except a dynamic symbol table that a kept REL/RELA links to
   * and that comes later in the input is hoisted in front of that reloc section. */
  int nk = 0;
  {
    char *emitted = (char *)calloc(shnum, 1);
    for (uint64_t i = 0; i < shnum; i++) {
      if (!S[i].keep || emitted[i]) continue;
      const Shdr *h = &S[i].h;
      if ((h->sh_type == SHT_REL || h->sh_type == SHT_RELA) && h->sh_link < shnum && h->sh_link > i &&
          S[h->sh_link].keep && !emitted[h->sh_link] &&
          (S[h->sh_link].h.sh_type == SHT_DYNSYM || S[h->sh_link].h.sh_type == SHT_SYMTAB)) {
        order[nk++] = (int)h->sh_link;
        emitted[h->sh_link] = 1;
      }
      order[nk++] = (int)i;
      emitted[i] = 1;
    }
    free(emitted);
  }
  for (int k = 0; k < nk; k++) S[order[k]].new_index = k;
  int new_shnum = nk + 1; /* + .shstrtab */

  /* R9: merge build-attribute notes before layout (their size feeds R4). */
  if (!(flags & LBO_FLAG_NO_MERGE_NOTES)) {
    for (uint64_t i = 1; i < shnum; i++) {
      if (!S[i].keep || S[i].h.sh_type != SHT_NOTE || (S[i].h.sh_flags & SHF_ALLOC)) continue;
      if (!has_prefix(S[i].name, ".gnu.build.attributes")) continue;
      uint64_t sz = S[i].h.sh_size;
      S[i].owned = (uint8_t *)malloc(sz * 2 + 16);
      int err = 0;
      S[i].new_size = merge_build_notes(S[i].data, sz, S[i].owned, &err);
      S[i].data = S[i].owned;
      if (err) { rc = LBO_BAD_NOTES; goto done; }
    }
  }

  /* R10/R11: lay out PT_LOADs.  Identity on linker-native files. */
  int new_phnum = 0;
  for (uint64_t i = 1; i < shnum; i++) {
    if (!S[i].keep || !(S[i].h.sh_flags & SHF_ALLOC)) continue;
    for (uint64_t j = 0; j < phnum; j++)
      if (P[j].p_type == PT_LOAD && sec_in_seg(&S[i].h, &P[j])) { S[i].seg = (int)j; break; }
    if (S[i].seg < 0) UNSUP();
    /* gate: file offset and address of a loaded section must move together ("lma adjusted" otherwise) */
    if (S[i].h.sh_type != SHT_NOBITS && S[i].h.sh_offset - P[S[i].seg].p_offset != S[i].h.sh_addr - P[S[i].seg].p_vaddr) UNSUP();
  }
  for (uint64_t j = 0; j < phnum; j++) {
    pkeep[j] = 1;
    if (P[j].p_type == PT_LOAD && P[j].p_offset != 0) {
      int members = 0;
      for (uint64_t i = 1; i < shnum; i++) members += (S[i].keep && S[i].seg == (int)j);
      if (!members) pkeep[j] = 0;
    }
    new_phnum += pkeep[j];
  }
  uint64_t cur = 64 + (uint64_t)new_phnum * 56;
  {
    uint64_t last_vaddr = 0;
    for (uint64_t j = 0; j < phnum; j++) {
      NP[j] = P[j];
      if (P[j].p_type != PT_LOAD || !pkeep[j]) continue;
      if (P[j].p_vaddr < last_vaddr) UNSUP();
      last_vaddr = P[j].p_vaddr;
      uint64_t new_off;
      int first = (P[j].p_offset == 0);
      int contents = 0;
      for (uint64_t i = 1; i < shnum; i++)
        if (S[i].keep && S[i].seg == (int)j && S[i].h.sh_type != SHT_NOBITS) contents = 1;
      if (first) new_off = 0;
      else {
        uint64_t al = P[j].p_align ? P[j].p_align : 1;
        new_off = cur + ((P[j].p_vaddr - cur) % al);
      }
      uint64_t off = first ? cur : new_off;
      uint64_t mem_end = P[j].p_vaddr + (first ? cur : 0);
      uint64_t file_end = off;
      int idx = 0;
      /* members are visited in input index order (== address order for well-formed inputs; the
       * R2 hoist changes the header order only, not the layout) */
      for (uint64_t i = 1; i < shnum; i++) {
        Sec *s = &S[i];
        if (!s->keep || s->seg != (int)j) continue;
        uint64_t want = new_off + (s->h.sh_addr - P[j].p_vaddr);
        if (s->h.sh_type != SHT_NOBITS) {
          if (want < off) UNSUP();
          off = want;
          s->new_off = off;
          off += s->new_size;
          file_end = off;
        } else {
          /* NOBITS: sh_offset only tracks the address when first in the segment */
          if (idx == 0) off = want;
          s->new_off = off;
        }
        if (!(s->h.sh_type == SHT_NOBITS && (s->h.sh_flags & SHF_TLS))) {
          uint64_t e = s->h.sh_addr + s->h.sh_size;
          if (e > mem_end) mem_end = e;
        }
        idx++;
      }
      NP[j].p_offset = new_off;
      if (!contents && !first) {
        uint64_t al = P[j].p_align > 0x1000 ? P[j].p_align : 0x1000;
        NP[j].p_offset = cur % al;
        NP[j].p_filesz = 0;
      } else {
        NP[j].p_filesz = file_end - new_off;
      }
      NP[j].p_memsz = mem_end - P[j].p_vaddr;
      pshift[j] = (int64_t)(NP[j].p_offset - P[j].p_offset);
      if (contents || first) cur = file_end;
    }
  }
  /* R12: non-LOAD program headers (elf.c assign_file_positions_for_non_load_sections). */
  for (uint64_t j = 0; j < phnum; j++) {
    uint32_t t = P[j].p_type;
    if (t == PT_LOAD || !pkeep[j]) continue;
    if (t == PT_PHDR) {
      NP[j].p_filesz = NP[j].p_memsz = (uint64_t)new_phnum * 56;
      continue;
    }
    /* members, in input order; the input headers decide membership */
    int64_t first = -1, last_bits = -1;
    for (uint64_t i = 1; i < shnum; i++) {
      if (!S[i].keep || !sec_in_seg(&in_hdr[i], &P[j])) continue;
      if (first < 0) first = (int64_t)i;
      if (S[i].h.sh_type != SHT_NOBITS) last_bits = (int64_t)i;
    }
    if (t == PT_GNU_STACK) { NP[j].p_offset = 0; NP[j].p_filesz = 0; continue; }
    if (first >= 0 && t != PT_GNU_RELRO && t != PT_TLS) {
      /* gate: a segment that carries sections must describe exactly their extent (BFD recomputes
       * offset/filesz/memsz from the sections; natural files already agree) */
      uint64_t aend = 0, fend = 0;
      int any_alloc = 0;
      for (uint64_t i = 1; i < shnum; i++) {
        if (!S[i].keep || !sec_in_seg(&in_hdr[i], &P[j])) continue;
        if (in_hdr[i].sh_flags & SHF_ALLOC) { any_alloc = 1; if (in_hdr[i].sh_addr + in_hdr[i].sh_size > aend) aend = in_hdr[i].sh_addr + in_hdr[i].sh_size; }
        if (in_hdr[i].sh_type != SHT_NOBITS && in_hdr[i].sh_offset + in_hdr[i].sh_size > fend) fend = in_hdr[i].sh_offset + in_hdr[i].sh_size;
      }
      if (any_alloc && (in_hdr[first].sh_addr != P[j].p_vaddr || aend - P[j].p_vaddr != P[j].p_memsz)) UNSUP();
      if (in_hdr[first].sh_type != SHT_NOBITS && in_hdr[first].sh_offset != P[j].p_offset) UNSUP();
      if (fend && fend - P[j].p_offset != P[j].p_filesz) UNSUP();
    }
    if (t == PT_GNU_RELRO) {
      int ok = 0;
      if (first >= 0) {
        uint64_t start = S[first].h.sh_addr, end = start + P[j].p_memsz;
        for (uint64_t l = 0; l < phnum && !ok; l++) {
          if (P[l].p_type != PT_LOAD || !pkeep[l]) continue;
          int64_t lf = -1, ll = -1;
          for (uint64_t i = 1; i < shnum; i++)
            if (S[i].keep && S[i].seg == (int)l) { if (lf < 0) lf = (int64_t)i; ll = (int64_t)i; }
          if (lf < 0) continue;
          uint64_t lend = S[ll].h.sh_addr +
                          ((S[ll].h.sh_type == SHT_NOBITS && (S[ll].h.sh_flags & SHF_TLS)) ? 0 : S[ll].h.sh_size);
          if (!(lend > start && S[lf].h.sh_addr < end)) continue;
          for (uint64_t i = 1; i < shnum; i++) {
            if (!S[i].keep || S[i].seg != (int)l) continue;
            if (S[i].h.sh_addr >= start && S[i].h.sh_addr < end && S[i].h.sh_size != 0) {
              NP[j].p_vaddr = S[i].h.sh_addr;
              NP[j].p_paddr = S[i].h.sh_addr + (P[l].p_paddr - P[l].p_vaddr);
              NP[j].p_offset = S[i].new_off;
              NP[j].p_memsz = end - NP[j].p_vaddr;
              NP[j].p_filesz = NP[j].p_memsz;
              if (NP[j].p_filesz > NP[l].p_vaddr + NP[l].p_filesz - NP[j].p_vaddr)
                NP[j].p_filesz = NP[l].p_vaddr + NP[l].p_filesz - NP[j].p_vaddr;
              ok = 1;
              break;
            }
          }
          break;
        }
      }
      if (!ok) memset(&NP[j], 0, sizeof(Phdr));
      continue;
    }
    if (first < 0) { NP[j].p_offset = 0; NP[j].p_filesz = 0; NP[j].p_memsz = 0; continue; }
    NP[j].p_offset = S[first].new_off;
    NP[j].p_filesz = 0;
    if (t == PT_TLS) {
      /* BFD recomputes PT_TLS p_memsz as the address extent of .tdata/.tbss (gold rounds the
       * input value up to the alignment; probe: -fuse-ld=gold fixture 0x40 -> 0x31) */
      uint64_t end = P[j].p_vaddr;
      for (uint64_t i = 1; i < shnum; i++)
        if (S[i].keep && sec_in_seg(&in_hdr[i], &P[j]) && S[i].h.sh_addr + S[i].h.sh_size > end) end = S[i].h.sh_addr + S[i].h.sh_size;
      NP[j].p_memsz = end - P[j].p_vaddr;
    }
    if (last_bits >= 0) {
      NP[j].p_filesz = S[last_bits].new_off - NP[j].p_offset + S[last_bits].new_size;
      if (t == PT_NOTE && (S[last_bits].h.sh_flags & SHF_ALLOC)) NP[j].p_memsz = NP[j].p_filesz;
    }
  }

  /* R4: non-alloc kept sections packed behind the last allocated byte. */
  for (int k = 1; k < nk; k++) {
    Sec *s = &S[order[k]];
    if (s->h.sh_flags & SHF_ALLOC) continue;
    uint64_t al = s->h.sh_addralign ? s->h.sh_addralign : 1;
    cur = align_up(cur, al);
    s->new_off = cur;
    if (s->h.sh_type != SHT_NOBITS) cur += s->new_size;
  }

  /* R6: names.  Insertion order: ".shstrtab" first (prep_headers), then sections in output order. */
  int nent = 0;
  ents[nent].s = ".shstrtab"; ents[nent].len = 9; ents[nent].order = 0; nent++;
  for (int k = 1; k < nk; k++) {
    const char *nm = S[order[k]].name;
    int dup = 0;
    for (int e = 0; e < nent; e++) if (strcmp(ents[e].s, nm) == 0) { dup = 1; break; }
    if (!dup && nm[0]) { ents[nent].s = nm; ents[nent].len = (int)strlen(nm); ents[nent].order = nent; nent++; }
  }
  uint32_t strsz = 0;
  strblob = build_strtab(ents, nent, &strsz);

  /* R5 */
  uint64_t shstr_off = cur;
  cur += strsz;
  uint64_t new_shoff = align_up(cur, 8);
  uint64_t total = new_shoff + (uint64_t)new_shnum * 64;

  if (total > n + ((uint64_t)65 << 20)) UNSUP();   /* wrapped address arithmetic on a corrupt file (see plan.cu) */
  out = (uint8_t *)calloc(total ? total : 1, 1);
  if (!out) { rc = LBO_NOMEM; goto done; }

  /* R8 */
  Ehdr neh = eh;
  neh.e_shoff = new_shoff;
  neh.e_shnum = (uint16_t)new_shnum;
  neh.e_shstrndx = (uint16_t)(new_shnum - 1);
  neh.e_phnum = (uint16_t)new_phnum;
  memcpy(out, &neh, 64);
  {
    uint64_t w = 64;
    for (uint64_t j = 0; j < phnum; j++)
      if (pkeep[j]) { memcpy(out + w, &NP[j], 56); w += 56; }
  }
  /* section contents */
  for (int k = 1; k < nk; k++) {
    Sec *s = &S[order[k]];
    if (s->h.sh_type == SHT_NOBITS || s->new_size == 0) continue;
    if (s->new_off + s->new_size > total) UNSUP();
    memcpy(out + s->new_off, s->data, s->new_size);
  }
  memcpy(out + shstr_off, strblob, strsz);

  /* R7 */
  for (int k = 1; k < nk; k++) {
    Sec *s = &S[order[k]];
    Shdr h = s->h;
    h.sh_name = 0;
    for (int e = 0; e < nent; e++) if (strcmp(ents[e].s, s->name) == 0) { h.sh_name = ents[e].off; break; }
    h.sh_offset = s->new_off;
    h.sh_size = s->new_size;
    if (h.sh_link && h.sh_link < shnum) h.sh_link = S[h.sh_link].keep ? (uint32_t)S[h.sh_link].new_index : 0;
    if ((h.sh_flags & SHF_INFO_LINK) && h.sh_info && h.sh_info < shnum)
      h.sh_info = S[h.sh_info].keep ? (uint32_t)S[h.sh_info].new_index : 0;
    if ((h.sh_type == SHT_REL || h.sh_type == SHT_RELA) && s->h.sh_link == 0) {
      /* assign_section_numbers(): an allocated reloc section without a symbol table gets .dynsym */
      for (int q = 1; q < nk; q++) if (strcmp(S[order[q]].name, ".dynsym") == 0) { h.sh_link = (uint32_t)q; break; }
    }
    if (h.sh_type == SHT_REL || h.sh_type == SHT_RELA) {
      /* BFD re-derives the section a dynamic reloc section applies to from its NAME
       * (elf.c: elf_get_reloc_section / _bfd_elf_plt_get_reloc_section): strip ".rel[a]",
       * and for ".plt" use ".got.plt" (else ".got"). */
      const char *nm = s->name;
      const char *t = NULL;
      if (has_prefix(nm, ".rela")) t = nm + 5;
      else if (has_prefix(nm, ".rel")) t = nm + 4;
      int target = -1;
      if (t && *t) {
        const char *alt = NULL;
        if (strcmp(t, ".plt") == 0) {
          for (int q = 1; q < nk; q++) if (strcmp(S[order[q]].name, ".got.plt") == 0) { target = q; break; }
          if (target < 0) alt = ".got";
          if (alt) for (int q = 1; q < nk; q++) if (strcmp(S[order[q]].name, alt) == 0) { target = q; break; }
        } else {
          for (int q = 1; q < nk; q++) if (strcmp(S[order[q]].name, t) == 0) { target = q; break; }
        }
      }
      /* SHF_INFO_LINK and sh_info exist in the output exactly when BFD finds the target section */
      h.sh_flags &= ~(uint64_t)SHF_INFO_LINK;
      h.sh_info = 0;
      if (target >= 0) { h.sh_info = (uint32_t)target; h.sh_flags |= SHF_INFO_LINK; }
    }
    if (h.sh_flags & (0x10 | 0x20)) h.sh_entsize &= 0xffffffffu;   /* SHF_MERGE/STRINGS: BFD carries entsize in an unsigned int */
    /* elf.c elf_fake_sections(): BFD recomputes sh_entsize for the section types it knows
     * (probe: doctored sh_entsize on each section of a gcc-built .so, binutils 2.42). */
    switch (h.sh_type) {
      case 14: case 15: case 16: case 19: h.sh_entsize = 8; break;  /* INIT/FINI/PREINIT_ARRAY, RELR */
      case 5: h.sh_entsize = 4; break;                              /* SHT_HASH (x86-64/aarch64) */
      case 6: h.sh_entsize = 16; break;                             /* SHT_DYNAMIC */
      case 0x6ffffff6: h.sh_entsize = 0; break;                     /* SHT_GNU_HASH, 64-bit */
      case 0x6ffffffd: case 0x6ffffffe: h.sh_entsize = 0; break;    /* GNU_verdef / GNU_verneed */
      default: break;
    }
    memcpy(out + new_shoff + (uint64_t)k * 64, &h, 64);
  }
  {
    Shdr h;
    memset(&h, 0, sizeof h);
    h.sh_name = ents[0].off;
    h.sh_type = SHT_STRTAB;
    h.sh_offset = shstr_off;
    h.sh_size = strsz;
    h.sh_addralign = 1;
    memcpy(out + new_shoff + (uint64_t)nk * 64, &h, 64);
  }
  *out_p = out;
  *out_n = total;
  out = NULL;

done:
  if (S) for (uint64_t i = 0; i < shnum; i++) free(S[i].owned);
  free(S); free(in_hdr); free(P); free(NP); free(pkeep); free(pshift); free(ents); free(order); free(strblob); free(out);
  return rc;
}

void lbo_free(uint8_t *p) { free(p); }

int lbo_last_unsupported_line(void) { return lbo_debug_line; }

const char *lbo_version(void) { return "strip_oracle restating GNU strip (Binutils 2.42) for ELF64-LE ET_DYN/ET_EXEC"; }

#ifdef LBO_MAIN
/* strip_oracle [-n] IN OUT : exit 0 on success, 10+rc for unsupported classes, 1 otherwise */
int main(int argc, char **argv) {
  unsigned flags = 0;
  int a = 1;
  if (argc > 1 && strcmp(argv[1], "-n") == 0) { flags |= LBO_FLAG_NO_MERGE_NOTES; a++; }
  if (argc - a != 2) { fprintf(stderr, "usage: strip_oracle [-n] IN OUT\n"); return 2; }
  FILE *f = fopen(argv[a], "rb");
  if (!f) { perror(argv[a]); return 1; }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t *buf = (uint8_t *)malloc(n ? (size_t)n : 1);
  if (fread(buf, 1, (size_t)n, f) != (size_t)n) { perror("read"); return 1; }
  fclose(f);
  uint8_t *out; uint64_t on;
  int rc = lbo_strip(buf, (uint64_t)n, &out, &on, flags);
  if (rc != 0) { fprintf(stderr, "strip_oracle: %s: rc=%d (line %d)\n", argv[a], rc, lbo_debug_line); return rc > 0 ? 10 + rc : 1; }
  f = fopen(argv[a + 1], "wb");
  if (!f) { perror(argv[a + 1]); return 1; }
  fwrite(out, 1, on, f);
  fclose(f);
  lbo_free(out);
  free(buf);
  return 0;
}
#endif
