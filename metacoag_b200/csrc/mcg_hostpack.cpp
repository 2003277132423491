// Host half of the contig ingest: ASCII bases -> the 2-bit packed store layout.
//
//   reference: feature_utils.py:18-35 (nt_bits / mer2bits): A0 C1 G2 T3 and ANY other
//   byte 0 (lower case and N included, feature_utils.py:21).
//
// Same output, bit for bit, as pack_ascii_kernel (mcg_scan.cu): base i of a contig in bits
// [2(i%16), +2) of 32-bit word i/16, every contig padded with zeros to a multiple of 64 bases
// (16 bytes).  This is format conversion for the PCIe link, not part of the scored
// arithmetic: mcg_load_contigs splits the contig blob between the copy engine (ASCII chunks,
// packed by the GPU kernel) and these routines (host threads pack, the 4x smaller result is
// copied), so that an upload costs max(link, host) instead of 8 bits per base on the link.
//
// Compiled by g++ (not nvcc): AVX-512BW / AVX2 + BMI2 variants picked at run time, scalar
// otherwise.
#include <immintrin.h>

#include <cstdint>
#include <cstring>

namespace {

// The packed block of 64 bases.  The output is written once and next read by the copy engine:
// a non-temporal store (when the block is 16-byte aligned, as it is in the pinned staging
// buffer) saves the read-for-ownership of the line, and the upload is bound by host DRAM
// traffic once enough threads pack.
__attribute__((target("sse2"))) inline void store16(uint8_t *dst, uint64_t w0, uint64_t w1) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst),
                         _mm_set_epi64x(static_cast<long long>(w1), static_cast<long long>(w0)));
    } else {
        memcpy(dst, &w0, 8);
        memcpy(dst + 8, &w1, 8);
    }
}

// 64 bases (mask = which of them exist) -> 16 bytes.
__attribute__((target("avx512bw,avx512vl,bmi2"))) inline void block_avx512(const uint8_t *src,
                                                                            uint64_t mask,
                                                                            uint8_t *dst) {
    const __m512i v = _mm512_maskz_loadu_epi8(mask, src);   // masked-off bytes are not touched
    const uint64_t mc = _mm512_cmpeq_epi8_mask(v, _mm512_set1_epi8('C'));
    const uint64_t mg = _mm512_cmpeq_epi8_mask(v, _mm512_set1_epi8('G'));
    const uint64_t mt = _mm512_cmpeq_epi8_mask(v, _mm512_set1_epi8('T'));
    const uint64_t lo = mc | mt, hi = mg | mt;
    const uint64_t w0 = _pdep_u64(lo, 0x5555555555555555ull) | _pdep_u64(hi, 0xAAAAAAAAAAAAAAAAull);
    const uint64_t w1 = _pdep_u64(lo >> 32, 0x5555555555555555ull) |
                        _pdep_u64(hi >> 32, 0xAAAAAAAAAAAAAAAAull);
    store16(dst, w0, w1);
}

__attribute__((target("avx512bw,avx512vl,bmi2"))) void contig_avx512(const uint8_t *src,
                                                                      int64_t len, uint8_t *dst) {
    const int64_t full = len / 64;
    for (int64_t i = 0; i < full; ++i) block_avx512(src + 64 * i, ~0ull, dst + 16 * i);
    const int rem = static_cast<int>(len - 64 * full);
    if (rem) block_avx512(src + 64 * full, (1ull << rem) - 1ull, dst + 16 * full);
}

__attribute__((target("avx2,bmi2"))) inline uint64_t half_avx2(const uint8_t *src) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
    const uint32_t mc = _mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('C')));
    const uint32_t mg = _mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('G')));
    const uint32_t mt = _mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('T')));
    const uint64_t lo = mc | mt, hi = mg | mt;
    return _pdep_u64(lo, 0x5555555555555555ull) | _pdep_u64(hi, 0xAAAAAAAAAAAAAAAAull);
}

inline uint32_t code_of(uint8_t c) { return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u; }

void tail_scalar(const uint8_t *src, int64_t len, uint8_t *dst) {
    // len < 64 bases into one zero-padded 16-byte block
    uint64_t w[2] = {0, 0};
    for (int64_t i = 0; i < len; ++i) w[i >> 5] |= static_cast<uint64_t>(code_of(src[i])) << (2 * (i & 31));
    memcpy(dst, w, 16);
}

__attribute__((target("avx2,bmi2"))) void contig_avx2(const uint8_t *src, int64_t len,
                                                      uint8_t *dst) {
    const int64_t full = len / 64;
    for (int64_t i = 0; i < full; ++i) {
        const uint64_t w0 = half_avx2(src + 64 * i), w1 = half_avx2(src + 64 * i + 32);
        store16(dst + 16 * i, w0, w1);
    }
    if (len > 64 * full) tail_scalar(src + 64 * full, len - 64 * full, dst + 16 * full);
}

__attribute__((target("avx512bw,avx512vl,bmi2"), noinline)) void tail_avx512(const uint8_t *src,
                                                                             int rem,
                                                                             uint8_t *dst) {
    block_avx512(src, (1ull << rem) - 1ull, dst);
}

// AVX2 body (measured faster than the 512-bit loop on the B200 hosts: 98 vs 81 GB/s with 16
// threads) with the masked 512-bit load for the partial block.
__attribute__((target("avx2,bmi2"))) void contig_avx2_tail512(const uint8_t *src, int64_t len,
                                                              uint8_t *dst) {
    const int64_t full = len / 64;
    for (int64_t i = 0; i < full; ++i) {
        const uint64_t w0 = half_avx2(src + 64 * i), w1 = half_avx2(src + 64 * i + 32);
        store16(dst + 16 * i, w0, w1);
    }
    if (len > 64 * full) tail_avx512(src + 64 * full, static_cast<int>(len - 64 * full), dst + 16 * full);
}

void contig_scalar(const uint8_t *src, int64_t len, uint8_t *dst) {
    const int64_t full = len / 64;
    for (int64_t i = 0; i < full; ++i) tail_scalar(src + 64 * i, 64, dst + 16 * i);
    if (len > 64 * full) tail_scalar(src + 64 * full, len - 64 * full, dst + 16 * full);
}

typedef void (*contig_fn)(const uint8_t *, int64_t, uint8_t *);

// level: -1 best available, 0 scalar, 1 AVX2, 2 AVX-512, 3 AVX2 body + AVX-512 tail; a level
// the CPU lacks falls back
contig_fn pick(int level) {
    __builtin_cpu_init();
    const bool bmi2 = __builtin_cpu_supports("bmi2");
    const bool has512 = bmi2 && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
    const bool has2 = bmi2 && __builtin_cpu_supports("avx2");
    if (level < 0) level = 3;
    if (level >= 3 && has512 && has2) return contig_avx2_tail512;
    if (level >= 2 && has512) return contig_avx512;
    if (level >= 1 && has2) return contig_avx2;
    return contig_scalar;
}

}  // namespace

// Pack contigs [c0, c1): contig c has offsets[c+1] - offsets[c] ASCII bases at bases +
// offsets[c]; its packed words start at packed + 4 * first_word[c * word_stride] bytes
// (first_word addressed with a byte stride so that it can point into an array of structs).
// level: -1 best available, 0 scalar, 1 AVX2, 2 AVX-512, 3 AVX2 + 512-bit tail (tests force
// the variants).
extern "C" void mcg_hostpack_range(const uint8_t *bases, const int64_t *offsets, int64_t c0,
                                   int64_t c1, const void *first_word, int64_t word_stride_bytes,
                                   uint8_t *packed, int level) {
    static contig_fn best = pick(-1);
    const contig_fn fn = level < 0 ? best : pick(level);
    const char *fw = static_cast<const char *>(first_word);
    for (int64_t c = c0; c < c1; ++c) {
        int64_t word;
        memcpy(&word, fw + c * word_stride_bytes, sizeof word);
        fn(bases + offsets[c], offsets[c + 1] - offsets[c], packed + 4 * word);
    }
    _mm_sfence();   // the non-temporal stores are visible before the chunk is reported packed
}

// the variant `level = -1` resolves to on this CPU
extern "C" int mcg_hostpack_level(void) {
    const contig_fn fn = pick(-1);
    return fn == contig_avx2_tail512 ? 3 : fn == contig_avx512 ? 2 : fn == contig_avx2 ? 1 : 0;
}

// ---- FASTA -> contig blob ---------------------------------------------------------------------
//   reference: feature_utils.py:131-138, 182-186 read the contigs with Bio.SeqIO.parse(path,
//   "fasta") and keep str(record.seq) and record.id.  BioPython's FASTA reader (Bio.SeqIO.FastaIO
//   SimpleFastaParser) does, on a text-mode handle with universal newlines:
//     - skip everything before the first line that starts with '>';
//     - title = line[1:].rstrip(); id = first whitespace-separated word of the title ("" if none);
//     - sequence = "".join(line.rstrip() for the following lines) with every ' ' and '\r' removed.
//   The two passes below produce exactly that, as one byte blob + offsets instead of a list of
//   Python strings: pass 1 sizes the outputs, pass 2 fills caller-owned
//   (pinned) buffers.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdlib>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

inline bool py_space(uint8_t c) {   // str.isspace() on the ASCII range
    return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31);
}

struct FastaOut {
    uint8_t *bases = nullptr;
    int64_t *offsets = nullptr;
    uint8_t *names = nullptr;
    int64_t *name_offsets = nullptr;
    int64_t n_records = 0, n_bases = 0, name_bytes = 0;
    int64_t cap_records = 0, cap_bases = 0, cap_names = 0;   // pass 2: sizes of the buffers
    bool overflow = false;
};

// One logical line [p, e) (terminator excluded).
template <bool FILL>
inline void fasta_line(const uint8_t *p, const uint8_t *e, bool *in_record, FastaOut *o) {
    if (p < e && *p == '>') {
        *in_record = true;
        const uint8_t *q = p + 1;
        while (e > q && py_space(e[-1])) --e;        // rstrip
        while (q < e && py_space(*q)) ++q;           // split(None, 1)[0]
        const uint8_t *w = q;
        while (w < e && !py_space(*w)) ++w;
        if (FILL) {
            if (o->n_records >= o->cap_records || o->name_bytes + (w - q) > o->cap_names) {
                o->overflow = true;
                return;
            }
            o->name_offsets[o->n_records] = o->name_bytes;
            memcpy(o->names + o->name_bytes, q, static_cast<size_t>(w - q));
            o->offsets[o->n_records] = o->n_bases;
        }
        o->name_bytes += w - q;
        o->n_records += 1;
        return;
    }
    if (!*in_record) return;
    while (e > p && py_space(e[-1])) --e;            // rstrip
    if (e <= p) return;
    if (FILL && o->overflow) return;
    if (memchr(p, ' ', static_cast<size_t>(e - p)) == nullptr) {
        if (FILL) {
            if (o->n_bases + (e - p) > o->cap_bases) { o->overflow = true; return; }
            memcpy(o->bases + o->n_bases, p, static_cast<size_t>(e - p));
        }
        o->n_bases += e - p;
    } else {
        for (; p < e; ++p)
            if (*p != ' ') {   // ('\r' never reaches here: it ends a line)
                if (FILL) {
                    if (o->n_bases >= o->cap_bases) { o->overflow = true; return; }
                    o->bases[o->n_bases] = *p;
                }
                o->n_bases += 1;
            }
    }
}

template <bool FILL> void fasta_pass(const uint8_t *data, size_t size, FastaOut *o) {
    bool in_record = false;
    const uint8_t *p = data, *end = data + size;
    while (p < end) {
        // universal newlines: '\n', "\r\n" and a lone '\r' all end a line
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(p, '\n', static_cast<size_t>(end - p)));
        const uint8_t *e = nl ? nl : end;
        const uint8_t *next = nl ? nl + 1 : end;
        const uint8_t *cr = static_cast<const uint8_t *>(memchr(p, '\r', static_cast<size_t>(e - p)));
        if (cr) {
            e = cr;
            next = cr + 1;
            if (next < end && *next == '\n') ++next;
        }
        fasta_line<FILL>(p, e, &in_record, o);
        p = next;
    }
}

struct Mapped {
    const uint8_t *data = nullptr;
    size_t size = 0;
    int fd = -1;
    bool open(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        size = static_cast<size_t>(st.st_size);
        if (size == 0) return true;
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, size, MADV_SEQUENTIAL);
        data = static_cast<const uint8_t *>(m);
        return true;
    }
    ~Mapped() {
        if (data) munmap(const_cast<uint8_t *>(data), size);
        if (fd >= 0) close(fd);
    }
};

}  // namespace

// Record-aligned byte ranges of the file, one per worker: a range starts at a '>' that opens a
// line (or at 0).  MCG_FASTA_RANGE_BYTES sets the target size (default 8 MB; the tests use a
// few bytes to put a boundary at every record).
static std::vector<size_t> fasta_ranges(const uint8_t *data, size_t size) {
    size_t target = 8u << 20;
    if (const char *env = getenv("MCG_FASTA_RANGE_BYTES"))
        if (atoll(env) > 0) target = static_cast<size_t>(atoll(env));
    size_t parts = size / target;
    const size_t hw = std::max(1u, std::thread::hardware_concurrency());
    if (parts > 4 * hw) parts = 4 * hw;
    std::vector<size_t> cuts{0};
    for (size_t i = 1; i < parts; ++i) {
        size_t p = std::max(cuts.back() + 1, size / parts * i);
        while (p < size && !(data[p] == '>' && (data[p - 1] == '\n' || data[p - 1] == '\r'))) ++p;
        if (p >= size) break;
        cuts.push_back(p);
    }
    cuts.push_back(size);
    return cuts;
}

// Pass 1 over every range in parallel: counts per range.
static void fasta_count(const uint8_t *data, const std::vector<size_t> &cuts,
                        std::vector<FastaOut> *per_range) {
    const size_t n = cuts.size() - 1;
    per_range->assign(n, FastaOut());
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < n;)
            fasta_pass<false>(data + cuts[i], cuts[i + 1] - cuts[i], &(*per_range)[i]);
    };
    const size_t workers = std::min<size_t>(n, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (size_t t = 1; t < workers; ++t) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
}

// Pass 1.  Returns 0, or -1 when the file cannot be read.
extern "C" int mcg_fasta_index(const char *path, int64_t *n_records, int64_t *n_bases,
                               int64_t *name_bytes) {
    Mapped m;
    if (!m.open(path)) return -1;
    const std::vector<size_t> cuts = fasta_ranges(m.data, m.size);
    std::vector<FastaOut> per_range;
    fasta_count(m.data, cuts, &per_range);
    *n_records = *n_bases = *name_bytes = 0;
    for (const FastaOut &o : per_range) {
        *n_records += o.n_records;
        *n_bases += o.n_bases;
        *name_bytes += o.name_bytes;
    }
    return 0;
}

// Pass 2 into buffers sized by pass 1 (offsets and name_offsets have n_records + 1 entries):
// the ranges are counted again (cheap next to the copy), prefix sums give every range its
// place in the outputs, and the ranges fill them in parallel.
// Returns 0, -1 when the file cannot be read, -2 when it changed between the passes.
extern "C" int mcg_fasta_load(const char *path, uint8_t *bases, int64_t *offsets, uint8_t *names,
                              int64_t *name_offsets, int64_t n_records, int64_t n_bases,
                              int64_t name_bytes) {
    Mapped m;
    if (!m.open(path)) return -1;
    const std::vector<size_t> cuts = fasta_ranges(m.data, m.size);
    std::vector<FastaOut> counted;
    fasta_count(m.data, cuts, &counted);
    const size_t n = cuts.size() - 1;
    std::vector<FastaOut> fill(n);
    int64_t rec = 0, nb = 0, nn = 0;
    for (size_t i = 0; i < n; ++i) {
        fill[i].n_records = rec;
        fill[i].n_bases = nb;
        fill[i].name_bytes = nn;
        rec += counted[i].n_records;
        nb += counted[i].n_bases;
        nn += counted[i].name_bytes;
        fill[i].cap_records = rec;      // a range must stay inside its own slice
        fill[i].cap_bases = nb;
        fill[i].cap_names = nn;
        fill[i].bases = bases;
        fill[i].offsets = offsets;
        fill[i].names = names;
        fill[i].name_offsets = name_offsets;
    }
    if (rec != n_records || nb != n_bases || nn != name_bytes) return -2;
    offsets[0] = 0;
    name_offsets[0] = 0;
    if (n_records == 0) return 0;
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < n;)
            fasta_pass<true>(m.data + cuts[i], cuts[i + 1] - cuts[i], &fill[i]);
    };
    const size_t workers = std::min<size_t>(n, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (size_t t = 1; t < workers; ++t) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
    for (size_t i = 0; i < n; ++i)
        if (fill[i].overflow || fill[i].n_records != fill[i].cap_records ||
            fill[i].n_bases != fill[i].cap_bases || fill[i].name_bytes != fill[i].cap_names)
            return -2;
    offsets[n_records] = n_bases;
    name_offsets[n_records] = name_bytes;
    return 0;
}

// Per-bin FASTA output (metacoag:1210-1251): the reference parses the contigs file a second
// time and writes ">{record.id}\n{record.seq}\n" for every binned record, in file order, to
// the file of its bin.  The staged blob already holds record.id and str(record.seq) of every
// record in file order (mcg_fasta_load), so the second parse is not needed: file_of_record[i]
// picks the output of record i (-1 = in no bin), every output is sized, gathered and written
// by one worker (outputs are independent, so they go in parallel).  Files are truncated like
// the reference's open(..., "w+"); an output without records is created empty.
// Returns 0, or -(1 + index) of the first output that could not be written; -1000000000 for a
// file index out of range.
extern "C" int mcg_write_bin_fastas(const uint8_t *bases, const int64_t *offsets,
                                    const uint8_t *names, const int64_t *name_offsets,
                                    int64_t n_records, const int32_t *file_of_record,
                                    const char *const *paths, int32_t n_files) {
    if (n_files <= 0) return 0;
    std::vector<std::vector<int64_t>> members(static_cast<size_t>(n_files));
    for (int64_t i = 0; i < n_records; ++i) {
        const int32_t f = file_of_record[i];
        if (f < 0) continue;
        if (f >= n_files) return -1000000000;
        members[static_cast<size_t>(f)].push_back(i);
    }
    std::atomic<int32_t> next(0);
    std::atomic<int32_t> failed(n_files);
    auto work = [&]() {
        std::vector<uint8_t> buf;
        for (int32_t f; (f = next.fetch_add(1)) < n_files;) {
            const std::vector<int64_t> &recs = members[static_cast<size_t>(f)];
            size_t bytes = 0;
            for (int64_t i : recs)
                bytes += static_cast<size_t>(name_offsets[i + 1] - name_offsets[i]) +
                         static_cast<size_t>(offsets[i + 1] - offsets[i]) + 3;
            buf.resize(bytes);
            uint8_t *p = buf.data();
            for (int64_t i : recs) {
                const size_t nn = static_cast<size_t>(name_offsets[i + 1] - name_offsets[i]);
                const size_t nb = static_cast<size_t>(offsets[i + 1] - offsets[i]);
                *p++ = '>';
                if (nn) memcpy(p, names + name_offsets[i], nn);
                p += nn;
                *p++ = '\n';
                if (nb) memcpy(p, bases + offsets[i], nb);
                p += nb;
                *p++ = '\n';
            }
            bool ok = false;
            const int fd = ::open(paths[f], O_WRONLY | O_CREAT | O_TRUNC, 0666);
            if (fd >= 0) {
                ok = true;
                size_t done = 0;
                while (done < bytes) {
                    const ssize_t w = ::write(fd, buf.data() + done, bytes - done);
                    if (w < 0) {
                        if (errno == EINTR) continue;
                        ok = false;
                        break;
                    }
                    done += static_cast<size_t>(w);
                }
                if (::close(fd) != 0) ok = false;
            }
            if (!ok) {
                int32_t cur = failed.load();
                while (f < cur && !failed.compare_exchange_weak(cur, f)) {
                }
            }
        }
    };
    const size_t workers = std::min<size_t>(static_cast<size_t>(n_files),
                                            std::max(1u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (size_t t = 1; t < workers; ++t) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
    const int32_t bad = failed.load();
    return bad < n_files ? -(1 + bad) : 0;
}
