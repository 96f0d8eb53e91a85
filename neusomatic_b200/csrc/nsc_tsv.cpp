// Candidate-TSV decoder (host side): the C++ restatement of the reference's
// neusomatic/python/dataloader.py:38-58 (extract_zlib, candidate_loader_tsv) and of the tag parsing in
// NeuSomaticDataset.__getitem__ (dataloader.py:267-274).  File format (generate_dataset.py:654-657,
// 1569-1581): one line per candidate,
//     "{i}\t1\t{tag}\t{base64(zlib(uint8[5,32,23]))}[\t{ann} ...]\n"
//     tag = "{chrom}.{pos}.{ref}.{alt}.{DEL|INS|NONE|SNP}.{center}.{len}.{tumor_cov}.{normal_cov}"
// The decoder mmaps the file, indexes the lines and inflates a range of candidates straight into the
// caller's (pinned) batch buffers in the layout nsc_score_host_u8 / nsc_forward_u8 consume.
#include <fcntl.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "neusomatic_cuda.h"
#include "nsc_inflate.h"

namespace nsc {
void set_error(const char* fmt, ...);
}

struct nsc_tsv {
  int fd = -1;
  const char* data = nullptr;
  size_t size = 0;
  std::vector<size_t> line_start;   // offsets of non-empty lines (+ sentinel = size)
};

namespace {

constexpr int kMatrixBytes = 5 * 32 * 23;

struct Field { const char* p; size_t n; };

// split on runs of whitespace (python str.split() semantics used by dataloader.py:44)
inline bool next_field(const char*& p, const char* end, Field* f) {
  while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
  if (p >= end) return false;
  const char* s = p;
  // the end of the field = the first separator (a line's range runs to the start of the next non-blank line, so it may hold
  // newlines of blank lines); memchr, because the image field is ~1.3 KB
  const char* e = (const char*)memchr(p, '\t', (size_t)(end - p));
  if (!e) e = end;
  const char* q = (const char*)memchr(p, ' ', (size_t)(e - p));
  if (q) e = q;
  q = (const char*)memchr(p, '\r', (size_t)(e - p));
  if (q) e = q;
  q = (const char*)memchr(p, '\n', (size_t)(e - p));
  if (q) e = q;
  p = e;
  f->p = s;
  f->n = (size_t)(p - s);
  return true;
}

struct B64 {
  int8_t t[256];
  B64() {
    memset(t, -1, sizeof t);
    const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; ++i) t[(unsigned char)a[i]] = (int8_t)i;
  }
};
const B64 kB64;

// returns decoded length or -1
inline int b64_decode(const char* s, size_t n, uint8_t* out, size_t cap) {
  while (n > 0 && s[n - 1] == '=') --n;
  size_t o = 0;
  size_t i = 0;
  // whole groups: four characters -> three bytes, one validity test per group
  const size_t groups = n / 4;
  if (groups * 3 > cap) return -1;
  for (size_t g = 0; g < groups; ++g, i += 4) {
    const int a = kB64.t[(unsigned char)s[i]], b = kB64.t[(unsigned char)s[i + 1]], c = kB64.t[(unsigned char)s[i + 2]],
              d = kB64.t[(unsigned char)s[i + 3]];
    if ((a | b | c | d) < 0) return -1;
    const uint32_t v = ((uint32_t)a << 18) | ((uint32_t)b << 12) | ((uint32_t)c << 6) | (uint32_t)d;
    out[o] = (uint8_t)(v >> 16);
    out[o + 1] = (uint8_t)(v >> 8);
    out[o + 2] = (uint8_t)v;
    o += 3;
  }
  uint32_t acc = 0;
  int bits = 0;
  for (; i < n; ++i) {
    const int v = kB64.t[(unsigned char)s[i]];
    if (v < 0) return -1;
    acc = (acc << 6) | (uint32_t)v;
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      if (o >= cap) return -1;
      out[o++] = (uint8_t)(acc >> bits);
    }
  }
  return (int)o;
}

inline bool parse_int(const char* p, size_t n, long* v) {
  if (n == 0 || n > 18) return false;
  char buf[24];
  memcpy(buf, p, n);
  buf[n] = 0;
  char* e;
  *v = strtol(buf, &e, 10);
  return *e == 0;
}

// Exact fast path of decimal -> double for plain "[-]digits[.digits]" with at most 15 significant digits and at most 22
// fraction digits (Clinger): the digits are an exact integer below 2^53 and 10^k (k <= 22) is an exact double, so ONE
// correctly rounded division gives the same double as strtod / Python's float().  Returns false for anything else
// (exponents, inf/nan, long mantissas): the caller falls back to strtod.
inline bool fast_decimal(const char* p, size_t n, double* out) {
  static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  size_t i = 0;
  bool neg = false;
  if (i < n && (p[i] == '-' || p[i] == '+')) { neg = p[i] == '-'; ++i; }
  uint64_t mant = 0;
  int digits = 0, frac = 0, seen = 0;
  bool dot = false;
  for (; i < n; ++i) {
    const char c = p[i];
    if (c >= '0' && c <= '9') {
      ++seen;
      if (mant != 0 || c != '0') ++digits;
      if (digits > 15) return false;
      mant = mant * 10 + (uint64_t)(c - '0');
      if (dot) ++frac;
    } else if (c == '.' && !dot) {
      dot = true;
    } else {
      return false;
    }
  }
  if (seen == 0 || frac > 22) return false;
  const double v = (double)mant / kPow10[frac];
  *out = neg ? -v : v;
  return true;
}

const char* kTypes[4] = {"DEL", "INS", "NONE", "SNP"};   // call.py:407

struct DecodeJob {
  const nsc_tsv* t;
  int64_t first, n;
  int num_anns;
  uint8_t* matrices;
  int32_t* meta;
  float* anns255;
  int32_t* labels;
  const int64_t* rows = nullptr;     // candidate indices instead of the range [first, first + n)
  bool images = true;                // inflate the image fields (false: tags / annotations only)
  uint8_t* text = nullptr;           // staging for the device decoder: the lines are copied here verbatim ...
  size_t text_base = 0;              // ... text[0] = file offset text_base ...
  uint32_t* spans = nullptr;         // ... and [n][2] = (offset in text, characters) of every image field
  std::vector<std::string>* tag_blocks = nullptr;   // per block of 64 candidates: the tags, each followed by '\n' (no directory part)
  int32_t* tag_lens = nullptr;       // [n][2] = (len(ref), len(alt)) of the tags
  uint64_t* tag_hashes = nullptr;    // [n] FNV-1a of the tags (distinct hashes prove distinct tags without a dictionary)
  std::atomic<int64_t> next{0};
  std::atomic<int> failed{0};
  std::string err;
};

void decode_range(DecodeJob* job) {
  std::vector<uint8_t> comp(16384);
  static const bool fast_inflate = getenv("NSC_TSV_ZLIB") == nullptr;      // NSC_TSV_ZLIB=1: zlib only (A/B, tests)
  std::vector<nsc::inflate::Tables> dyn_store(1);
  nsc::inflate::Tables& dyn = dyn_store[0];
  uint8_t scratch[kMatrixBytes + 64];
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit(&zs) != Z_OK) { job->failed = 1; return; }
  for (;;) {
    const int64_t k0 = job->next.fetch_add(64);
    if (k0 >= job->n || job->failed.load()) break;
    const int64_t k1 = std::min<int64_t>(job->n, k0 + 64);
    for (int64_t k = k0; k < k1; ++k) {
      const size_t li = (size_t)(job->rows ? job->rows[k] : job->first + k);
      const char* p = job->t->data + job->t->line_start[li];
      const char* end = job->t->data + job->t->line_start[li + 1];
      while (end > p && (end[-1] == '\n' || end[-1] == '\r')) --end;
      Field f[4];
      bool ok = true;
      for (int i = 0; i < 4 && ok; ++i) ok = next_field(p, end, &f[i]);
      char msg[160] = "";
      if (!ok) snprintf(msg, sizeof msg, "line %zu: fewer than 4 fields", li);
      // tag: the last five dot-separated fields are type.center.length.tumor_cov.normal_cov (dataloader.py:267-274)
      long center = 0, length = 0, tcov = 0, ncov = 0;
      int type_idx = -1;
      if (ok) {
        const char* tb = f[2].p;
        const char* te = tb + f[2].n;
        for (const char* q = te; q > tb; --q)
          if (q[-1] == '/') { tb = q; break; }        // path_.split("/")[-1]
        const char* dots[5];
        int nd = 0;
        for (const char* q = te; q > tb && nd < 5; --q)
          if (q[-1] == '.') dots[nd++] = q - 1;
        // the reference unpacks exactly 9 fields
        int total_dots = 0;
        for (const char* q = tb; q < te; ++q) total_dots += *q == '.';
        if (nd < 5 || total_dots != 8) { ok = false; snprintf(msg, sizeof msg, "line %zu: tag does not have 9 fields", li); }
        if (ok) {
          ok = parse_int(dots[0] + 1, (size_t)(te - dots[0] - 1), &ncov) && parse_int(dots[1] + 1, (size_t)(dots[0] - dots[1] - 1), &tcov) &&
               parse_int(dots[2] + 1, (size_t)(dots[1] - dots[2] - 1), &length) && parse_int(dots[3] + 1, (size_t)(dots[2] - dots[3] - 1), &center);
          const size_t tn = (size_t)(dots[3] - dots[4] - 1);
          for (int i = 0; i < 4; ++i)
            if (strlen(kTypes[i]) == tn && memcmp(kTypes[i], dots[4] + 1, tn) == 0) type_idx = i;
          if (!ok) snprintf(msg, sizeof msg, "line %zu: non-integer tag field", li);
          // type_class_dict[tag.split(".")[4]] raises KeyError in the reference (dataloader.py:55)
          else if (type_idx < 0) { ok = false; snprintf(msg, sizeof msg, "line %zu: unknown variant type in tag", li); }
        }
        if (ok && job->tag_blocks) {
          // the dictionaries' key and what pred_vcf_records_none branches on (call.py:343-350): chrom.pos.REF.ALT.type...
          (*job->tag_blocks)[(size_t)(k0 / 64)].append(tb, (size_t)(te - tb)).push_back('\n');
          const char* d1 = (const char*)memchr(tb, '.', (size_t)(te - tb));
          const char* d2 = (const char*)memchr(d1 + 1, '.', (size_t)(te - d1 - 1));
          const char* d3 = (const char*)memchr(d2 + 1, '.', (size_t)(te - d2 - 1));
          job->tag_lens[2 * k] = (int32_t)(d3 - d2 - 1);
          job->tag_lens[2 * k + 1] = (int32_t)(dots[4] - d3 - 1);
          if (job->tag_hashes) {
            uint64_t h = 14695981039346656037ull;
            for (const char* q = tb; q < te; ++q) h = (h ^ (uint8_t)*q) * 1099511628211ull;
            job->tag_hashes[k] = h;
          }
        }
      }
      if (ok && job->text) {
        // staging: the whole line travels (one copy, no scatter); the device decoder finds the image field through its span
        const size_t l0 = job->t->line_start[li], l1 = job->t->line_start[li + 1];
        memcpy(job->text + (l0 - job->text_base), job->t->data + l0, l1 - l0);
        job->spans[2 * k] = (uint32_t)((size_t)(f[3].p - job->t->data) - job->text_base);
        job->spans[2 * k + 1] = (uint32_t)f[3].n;
      }
      // image: base64 -> zlib -> uint8[5,32,23]  (dataloader.py:38-39)
      if (ok && job->images) {
        const size_t need = f[3].n * 3 / 4 + 4 + 16;          // + the read-ahead slack of the word-wise bit buffer
        if (comp.size() < need) comp.resize(need);
        const int clen = b64_decode(f[3].p, f[3].n, comp.data(), comp.size() - 16);
        if (clen < 0) { ok = false; snprintf(msg, sizeof msg, "line %zu: invalid base64", li); }
        // fast path (nsc_inflate.h) into a private buffer with slack; anything it does not take is left to zlib, whose
        // verdict on a malformed stream is the reference's (dataloader.py:38-39 zlib.decompress)
        if (ok && fast_inflate && nsc::inflate::zlib_inflate_exact(comp.data(), (size_t)clen, scratch, kMatrixBytes, &dyn)) {
          memcpy(job->matrices + (size_t)k * kMatrixBytes, scratch, kMatrixBytes);
        } else if (ok) {
          inflateReset(&zs);
          zs.next_in = comp.data();
          zs.avail_in = (uInt)clen;
          zs.next_out = job->matrices + (size_t)k * kMatrixBytes;
          zs.avail_out = kMatrixBytes;
          const int rc = inflate(&zs, Z_FINISH);
          if (rc != Z_STREAM_END || zs.total_out != (uLong)kMatrixBytes) {
            ok = false;
            snprintf(msg, sizeof msg, "line %zu: zlib stream does not hold %d bytes", li, kMatrixBytes);
          }
        }
      }
      if (ok && job->meta) {
        job->meta[k * 3 + 0] = (int32_t)center;
        job->meta[k * 3 + 1] = (int32_t)tcov;
        job->meta[k * 3 + 2] = (int32_t)ncov;
        if (job->labels) { job->labels[k * 2 + 0] = type_idx; job->labels[k * 2 + 1] = (int32_t)(length < 3 ? length : 3); }   // varlen_label = min(length, 3), dataloader.py:415
        // annotations: float(text) * 255.0 in float64, cast to fp32 (dataloader.py:395-396, then .float())
        int na = 0;
        Field a;
        while (next_field(p, end, &a)) {
          if (na < job->num_anns && job->anns255) {
            double v;
            if (!fast_decimal(a.p, a.n, &v)) {      // exponents, long mantissas, inf / nan: the C library
              char buf[64];
              const size_t m = a.n < 63 ? a.n : 63;
              memcpy(buf, a.p, m);
              buf[m] = 0;
              char* e;
              v = strtod(buf, &e);
              if (*e != 0) { ok = false; snprintf(msg, sizeof msg, "line %zu: bad annotation '%s'", li, buf); break; }
            }
            job->anns255[(size_t)k * job->num_anns + na] = (float)(v * 255.0);
          }
          ++na;
        }
        if (ok && na != job->num_anns) { ok = false; snprintf(msg, sizeof msg, "line %zu: %d annotations, expected %d", li, na, job->num_anns); }
      }
      if (!ok) {
        if (job->failed.exchange(1) == 0) job->err = msg;
        break;
      }
    }
  }
  inflateEnd(&zs);
}

}  // namespace

namespace nsc {
// One image field on the host, zlib's verdict (what the device decoder's declined candidates get): base64 -> zlib -> 3680 bytes
bool tsv_inflate_field(const char* p, size_t n, uint8_t* out, char* msg, size_t msg_cap) {
  std::vector<uint8_t> comp(n * 3 / 4 + 4 + 16);
  const int clen = b64_decode(p, n, comp.data(), comp.size() - 16);
  if (clen < 0) { snprintf(msg, msg_cap, "invalid base64"); return false; }
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit(&zs) != Z_OK) { snprintf(msg, msg_cap, "inflateInit failed"); return false; }
  std::vector<uint8_t> tmp(kMatrixBytes + 8);
  zs.next_in = comp.data();
  zs.avail_in = (uInt)clen;
  zs.next_out = tmp.data();
  zs.avail_out = kMatrixBytes;
  const int rc = ::inflate(&zs, Z_FINISH);
  const bool ok = rc == Z_STREAM_END && zs.total_out == (uLong)kMatrixBytes;
  inflateEnd(&zs);
  if (!ok) { snprintf(msg, msg_cap, "zlib stream does not hold %d bytes", kMatrixBytes); return false; }
  memcpy(out, tmp.data(), kMatrixBytes);
  return true;
}
}  // namespace nsc

namespace {
int run_job(DecodeJob& job, int64_t n, int num_threads) {
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min<int>(nt, (int)((n + 63) / 64)));
  std::vector<std::thread> th;
  for (int i = 1; i < nt; ++i) th.emplace_back(decode_range, &job);
  decode_range(&job);
  for (auto& x : th) x.join();
  if (job.failed.load()) { nsc::set_error("%s", job.err.empty() ? "TSV decode failed" : job.err.c_str()); return NSC_E_INVALID; }
  return NSC_OK;
}
}  // namespace

extern "C" {

int nsc_tsv_open(nsc_tsv** out, const char* path) {
  if (!out || !path) { nsc::set_error("NULL argument"); return NSC_E_INVALID; }
  *out = nullptr;
  int fd = open(path, O_RDONLY);
  if (fd < 0) { nsc::set_error("cannot open '%s'", path); return NSC_E_INVALID; }
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); nsc::set_error("fstat failed on '%s'", path); return NSC_E_INVALID; }
  nsc_tsv* t = new nsc_tsv();
  t->fd = fd;
  t->size = (size_t)st.st_size;
  if (t->size > 0) {
    void* p = mmap(nullptr, t->size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) { close(fd); delete t; nsc::set_error("mmap failed on '%s'", path); return NSC_E_NOMEM; }
    t->data = (const char*)p;
    madvise(p, t->size, MADV_SEQUENTIAL);
  }
  // line index: the file is cut into pieces at arbitrary offsets, each piece's thread lists the non-blank lines that START in
  // it (a line starts at 0 or after a '\n'), and the lists are joined in order
  const size_t piece = (size_t)32 << 20;
  const int np = (int)std::max<size_t>(1, (t->size + piece - 1) / piece);
  std::vector<std::vector<size_t>> found((size_t)np);
  auto scan = [&](int pi) {
    const size_t lo = (size_t)pi * piece, hi = std::min(t->size, lo + piece);
    size_t pos = lo;
    if (lo > 0 && t->data[lo - 1] != '\n') {          // inside a line that started in an earlier piece
      const char* nl = (const char*)memchr(t->data + lo, '\n', t->size - lo);
      pos = nl ? (size_t)(nl - t->data) + 1 : t->size;
    }
    std::vector<size_t>& v = found[(size_t)pi];
    while (pos < hi) {
      const char* nl = (const char*)memchr(t->data + pos, '\n', t->size - pos);
      const size_t e = nl ? (size_t)(nl - t->data) + 1 : t->size;
      bool blank = true;
      for (size_t i = pos; i < e && blank; ++i) blank = t->data[i] == '\n' || t->data[i] == '\r' || t->data[i] == ' ' || t->data[i] == '\t';
      if (!blank) v.push_back(pos);
      pos = e;
    }
  };
  {
    const int nt = std::max(1, std::min<int>(np, (int)std::thread::hardware_concurrency()));
    std::atomic<int> next{0};
    auto worker = [&]() { for (int pi = next.fetch_add(1); pi < np; pi = next.fetch_add(1)) scan(pi); };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(worker);
    worker();
    for (auto& x : th) x.join();
  }
  size_t total = 0;
  for (auto& v : found) total += v.size();
  t->line_start.reserve(total + 1);
  for (auto& v : found) t->line_start.insert(t->line_start.end(), v.begin(), v.end());
  t->line_start.push_back(t->size);
  *out = t;
  return NSC_OK;
}

int64_t nsc_tsv_count(const nsc_tsv* t) { return t ? (int64_t)t->line_start.size() - 1 : -1; }

int nsc_tsv_decode(nsc_tsv* t, int64_t first, int64_t n, int num_anns, uint8_t* matrices, int32_t* meta, float* anns255,
                   int32_t* labels, int num_threads) {
  if (!t || first < 0 || n < 0 || first + n > nsc_tsv_count(t)) { nsc::set_error("candidate range out of bounds"); return NSC_E_INVALID; }
  if (n == 0) return NSC_OK;
  if (!matrices || !meta || num_anns < 0 || (num_anns > 0 && !anns255)) { nsc::set_error("NULL output buffer"); return NSC_E_INVALID; }
  DecodeJob job;
  job.t = t; job.first = first; job.n = n; job.num_anns = num_anns;
  job.matrices = matrices; job.meta = meta; job.anns255 = anns255; job.labels = labels;
  return run_job(job, n, num_threads);
}

int nsc_tsv_decode_rows(nsc_tsv* t, const int64_t* rows, int64_t n, uint8_t* matrices, int num_threads) {
  if (!t || n < 0 || (n > 0 && (!rows || !matrices))) { nsc::set_error("bad argument"); return NSC_E_INVALID; }
  const int64_t count = nsc_tsv_count(t);
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= count) { nsc::set_error("candidate index %lld out of bounds", (long long)rows[i]); return NSC_E_INVALID; }
  if (n == 0) return NSC_OK;
  DecodeJob job;
  job.t = t; job.first = 0; job.n = n; job.num_anns = 0;
  job.matrices = matrices; job.meta = nullptr; job.anns255 = nullptr; job.labels = nullptr;
  job.rows = rows;
  return run_job(job, n, num_threads);
}

int64_t nsc_tsv_range_bytes(const nsc_tsv* t, int64_t first, int64_t n) {
  if (!t || first < 0 || n < 0 || first + n > nsc_tsv_count(t)) return -1;
  return (int64_t)(t->line_start[(size_t)(first + n)] - t->line_start[(size_t)first]);
}

int nsc_tsv_stage(nsc_tsv* t, int64_t first, int64_t n, int num_anns, uint8_t* text, int64_t text_cap, uint32_t* spans, int32_t* meta,
                  float* anns255, int32_t* labels, int num_threads, char* tag_buf, int64_t tag_cap, int64_t* tag_needed, int32_t* tag_lens,
                  uint64_t* tag_hashes) {
  if (tag_buf != nullptr && (tag_needed == nullptr || tag_lens == nullptr)) { nsc::set_error("tag_buf without tag_needed / tag_lens"); return NSC_E_INVALID; }
  if (tag_needed) *tag_needed = 0;
  if (!t || first < 0 || n < 0 || first + n > nsc_tsv_count(t)) { nsc::set_error("candidate range out of bounds"); return NSC_E_INVALID; }
  if (n == 0) return NSC_OK;
  if (!text || !spans || !meta || num_anns < 0 || (num_anns > 0 && !anns255)) { nsc::set_error("NULL output buffer"); return NSC_E_INVALID; }
  const int64_t bytes = nsc_tsv_range_bytes(t, first, n);
  if (bytes > text_cap || bytes > 0xffffffffll) { nsc::set_error("text buffer too small: %lld of %lld bytes", (long long)text_cap, (long long)bytes); return NSC_E_INVALID; }
  DecodeJob job;
  job.t = t; job.first = first; job.n = n; job.num_anns = num_anns;
  job.matrices = nullptr; job.meta = meta; job.anns255 = anns255; job.labels = labels;
  job.images = false;
  job.text = text;
  job.text_base = t->line_start[(size_t)first];
  job.spans = spans;
  std::vector<std::string> blocks;
  if (tag_buf != nullptr) {
    blocks.resize((size_t)((n + 63) / 64));
    job.tag_blocks = &blocks;
    job.tag_lens = tag_lens;
    job.tag_hashes = tag_hashes;
  }
  const int rc = run_job(job, n, num_threads);
  if (rc != NSC_OK || tag_buf == nullptr) return rc;
  int64_t o = 0;
  for (const std::string& b : blocks) {
    if (o + (int64_t)b.size() <= tag_cap) memcpy(tag_buf + o, b.data(), b.size());
    o += (int64_t)b.size();
  }
  *tag_needed = o;
  if (o > tag_cap) { nsc::set_error("tag buffer too small: %lld of %lld bytes", (long long)tag_cap, (long long)o); return NSC_E_INVALID; }
  return NSC_OK;
}

int nsc_tsv_tag(const nsc_tsv* t, int64_t i, const char** tag, int32_t* len) {
  if (!t || !tag || !len || i < 0 || i >= nsc_tsv_count(t)) { nsc::set_error("bad argument"); return NSC_E_INVALID; }
  const char* p = t->data + t->line_start[(size_t)i];
  const char* end = t->data + t->line_start[(size_t)i + 1];
  Field f;
  for (int k = 0; k < 3; ++k)
    if (!next_field(p, end, &f)) { nsc::set_error("line %lld: no tag field", (long long)i); return NSC_E_INVALID; }
  *tag = f.p;
  *len = (int32_t)f.n;
  return NSC_OK;
}

int nsc_tsv_tags(const nsc_tsv* t, int64_t first, int64_t n, char* buf, int64_t cap, int64_t* needed) {
  if (!t || !needed || first < 0 || n < 0 || first + n > nsc_tsv_count(t)) { nsc::set_error("bad argument"); return NSC_E_INVALID; }
  int64_t o = 0;
  for (int64_t i = first; i < first + n; ++i) {
    const char* p = t->data + t->line_start[(size_t)i];
    const char* end = t->data + t->line_start[(size_t)i + 1];
    Field f;
    for (int k = 0; k < 3; ++k)
      if (!next_field(p, end, &f)) { nsc::set_error("line %lld: no tag field", (long long)i); return NSC_E_INVALID; }
    if (buf != nullptr && o + (int64_t)f.n + 1 <= cap) {
      memcpy(buf + o, f.p, f.n);
      buf[o + f.n] = '\n';
    }
    o += (int64_t)f.n + 1;
  }
  *needed = o;
  if (buf != nullptr && o > cap) { nsc::set_error("tag buffer too small: %lld of %lld bytes", (long long)cap, (long long)o); return NSC_E_INVALID; }
  return NSC_OK;
}

int nsc_tsv_tags_lens(const nsc_tsv* t, int64_t first, int64_t n, char* buf, int64_t cap, int64_t* needed, int32_t* lens) {
  if (!t || !needed || !lens || first < 0 || n < 0 || first + n > nsc_tsv_count(t)) { nsc::set_error("bad argument"); return NSC_E_INVALID; }
  int64_t o = 0;
  for (int64_t i = first; i < first + n; ++i) {
    const char* p = t->data + t->line_start[(size_t)i];
    const char* end = t->data + t->line_start[(size_t)i + 1];
    Field f;
    for (int k = 0; k < 3; ++k)
      if (!next_field(p, end, &f)) { nsc::set_error("line %lld: no tag field", (long long)i); return NSC_E_INVALID; }
    const char* tb = f.p;
    const char* te = f.p + f.n;
    for (const char* q = te; q > tb; --q)
      if (q[-1] == '/') { tb = q; break; }          // path_.split("/")[-1]
    const int64_t tn = (int64_t)(te - tb);
    if (buf != nullptr && o + tn + 1 <= cap) {
      memcpy(buf + o, tb, (size_t)tn);
      buf[o + tn] = '\n';
    }
    o += tn + 1;
    int field = 0;
    const char* fs = tb;
    int32_t rl = -1, al = -1;
    for (const char* q = tb; q <= te; ++q) {
      if (q == te || *q == '.') {
        if (field == 2) rl = (int32_t)(q - fs);
        if (field == 3) { al = (int32_t)(q - fs); break; }
        ++field;
        fs = q + 1;
      }
    }
    if (rl < 0 || al < 0) { nsc::set_error("line %lld: tag does not have 9 fields", (long long)i); return NSC_E_INVALID; }
    lens[(i - first) * 2] = rl;
    lens[(i - first) * 2 + 1] = al;
  }
  *needed = o;
  if (buf != nullptr && o > cap) { nsc::set_error("tag buffer too small: %lld of %lld bytes", (long long)cap, (long long)o); return NSC_E_INVALID; }
  return NSC_OK;
}

int nsc_tsv_allele_lens(const nsc_tsv* t, int64_t first, int64_t n, int32_t* lens) {
  if (!t || !lens || first < 0 || n < 0 || first + n > nsc_tsv_count(t)) { nsc::set_error("bad argument"); return NSC_E_INVALID; }
  for (int64_t i = first; i < first + n; ++i) {
    const char* p = t->data + t->line_start[(size_t)i];
    const char* end = t->data + t->line_start[(size_t)i + 1];
    Field f;
    for (int k = 0; k < 3; ++k)
      if (!next_field(p, end, &f)) { nsc::set_error("line %lld: no tag field", (long long)i); return NSC_E_INVALID; }
    const char* tb = f.p;
    const char* te = f.p + f.n;
    for (const char* q = te; q > tb; --q)
      if (q[-1] == '/') { tb = q; break; }          // path_.split("/")[-1]
    // chrom.pos.ref.alt.type...: lengths of the third and fourth dot-separated fields
    int field = 0;
    const char* fs = tb;
    int32_t rl = -1, al = -1;
    for (const char* q = tb; q <= te; ++q) {
      if (q == te || *q == '.') {
        if (field == 2) rl = (int32_t)(q - fs);
        if (field == 3) { al = (int32_t)(q - fs); break; }
        ++field;
        fs = q + 1;
      }
    }
    if (rl < 0 || al < 0) { nsc::set_error("line %lld: tag does not have 9 fields", (long long)i); return NSC_E_INVALID; }
    lens[(i - first) * 2] = rl;
    lens[(i - first) * 2 + 1] = al;
  }
  return NSC_OK;
}

int nsc_debug_inflate(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t out_len, int use_zlib) {
  if (!in || !out || in_len < 0 || out_len < 0) { nsc::set_error("bad argument"); return NSC_E_INVALID; }
  std::vector<uint8_t> src((size_t)in_len + 16, 0), dst((size_t)out_len + 64);
  memcpy(src.data(), in, (size_t)in_len);
  if (!use_zlib) {
    std::vector<nsc::inflate::Tables> t(1);
    if (!nsc::inflate::zlib_inflate_exact(src.data(), (size_t)in_len, dst.data(), (size_t)out_len, &t[0])) {
      nsc::set_error("fast inflate declined the stream");
      return NSC_E_STATE;
    }
  } else {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) { nsc::set_error("inflateInit failed"); return NSC_E_NOMEM; }
    zs.next_in = src.data();
    zs.avail_in = (uInt)in_len;
    zs.next_out = dst.data();
    zs.avail_out = (uInt)out_len;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && zs.total_out == (uLong)out_len && zs.avail_in == 0;
    inflateEnd(&zs);
    if (!ok) { nsc::set_error("zlib: stream does not hold %lld bytes", (long long)out_len); return NSC_E_INVALID; }
  }
  memcpy(out, dst.data(), (size_t)out_len);
  return NSC_OK;
}

int nsc_tsv_close(nsc_tsv* t) {
  if (!t) return NSC_OK;
  if (t->data) munmap((void*)t->data, t->size);
  if (t->fd >= 0) close(t->fd);
  delete t;
  return NSC_OK;
}

}  // extern "C"
