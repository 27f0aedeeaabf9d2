// Host-side corpus ingestion: text -> CSR, the native counterpart of
// VariationalBayes.parse_data (reference variational_bayes.py:98-130).  Plain C++ (no device code;
// it lives in a .cu file only so that the one nvcc recipe builds the whole library).
//
// Semantics reproduced exactly (token indexing must be bit-exact):
//   * a document is one line; tokens are separated by ASCII whitespace (str.split() of an ASCII line);
//   * tokens missing from the vocabulary are skipped (:107-108);
//   * per document: distinct type ids in FIRST-SEEN order (dict insertion order under Python 3)
//     with their counts (:110-113, :119-120);
//   * documents without a single in-vocabulary token are dropped (:115-117).
// Documents are parsed in parallel by host threads (one contiguous range of lines each) and
// concatenated in order.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pylda_b200.h"

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }

inline uint64_t fnv1a(const char* s, size_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; ++i) {
        h ^= (unsigned char)s[i];
        h *= 1099511628211ULL;
    }
    return h;
}

// open-addressing map: vocabulary word -> type id
struct Vocab {
    std::vector<int32_t> slot;      // -1 = empty, else word index
    std::vector<const char*> ptr;
    std::vector<uint32_t> len;
    uint64_t mask = 0;
    void build(const char* buf, int64_t n) {
        const char* p = buf;
        const char* end = buf + n;
        while (p < end) {
            const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
            if (!q) q = end;
            ptr.push_back(p);
            len.push_back((uint32_t)(q - p));
            p = q + 1;
        }
        uint64_t cap = 16;
        while (cap < 2 * ptr.size() + 1) cap <<= 1;
        mask = cap - 1;
        slot.assign((size_t)cap, -1);
        for (size_t i = 0; i < ptr.size(); ++i) {
            uint64_t h = fnv1a(ptr[i], len[i]) & mask;
            while (slot[(size_t)h] >= 0) h = (h + 1) & mask;      // duplicate words: first id wins at lookup
            slot[(size_t)h] = (int32_t)i;
        }
    }
    int32_t find(const char* s, size_t n) const {
        uint64_t h = fnv1a(s, n) & mask;
        while (true) {
            const int32_t i = slot[(size_t)h];
            if (i < 0) return -1;
            if (len[(size_t)i] == n && memcmp(ptr[(size_t)i], s, n) == 0) return i;
            h = (h + 1) & mask;
        }
    }
};

struct Part {
    std::vector<int64_t> doc_nnz;   // per kept document
    std::vector<int32_t> ids, cts;
    int64_t dropped = 0;
};

void parse_range(const Vocab& vocab, const char* text, const std::vector<int64_t>& line_start, int64_t lo, int64_t hi,
                 Part* out) {
    const size_t V = vocab.ptr.size();
    std::vector<int64_t> stamp(V, -1);      // last document that saw the type
    std::vector<int32_t> pos(V, 0);         // its position in that document's list
    for (int64_t d = lo; d < hi; ++d) {
        const char* p = text + line_start[(size_t)d];
        const char* end = text + line_start[(size_t)d + 1];
        const size_t first = out->ids.size();
        while (p < end) {
            while (p < end && is_space((unsigned char)*p)) ++p;
            const char* q = p;
            while (q < end && !is_space((unsigned char)*q)) ++q;
            if (q > p) {
                const int32_t id = vocab.find(p, (size_t)(q - p));
                if (id >= 0) {
                    if (stamp[(size_t)id] != d) {
                        stamp[(size_t)id] = d;
                        pos[(size_t)id] = (int32_t)(out->ids.size() - first);
                        out->ids.push_back(id);
                        out->cts.push_back(1);
                    } else {
                        out->cts[first + (size_t)pos[(size_t)id]] += 1;
                    }
                }
            }
            p = q;
        }
        const int64_t n = (int64_t)(out->ids.size() - first);
        if (n == 0) out->dropped++;
        else out->doc_nnz.push_back(n);
    }
}

}  // namespace

struct pylda_parsed {
    std::vector<int64_t> row_ptr;
    std::vector<int32_t> ids, cts;
    int64_t dropped = 0;
};

extern "C" {

int pylda_parse_corpus(const char* text, int64_t text_len, const char* vocab, int64_t vocab_len, int n_threads,
                       pylda_parsed** out) {
    if (!out || text_len < 0 || vocab_len < 0 || (text_len > 0 && !text) || (vocab_len > 0 && !vocab)) return 1;
    *out = nullptr;
    Vocab voc;
    voc.build(vocab, vocab_len);
    // line index (a trailing newline does not start another document)
    std::vector<int64_t> line_start;
    {
        const char* p = text;
        const char* end = text + text_len;
        while (p < end) {
            line_start.push_back((int64_t)(p - text));
            const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
            p = q ? q + 1 : end;
        }
        line_start.push_back(text_len);
    }
    const int64_t L = (int64_t)line_start.size() - 1;
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(T, 64), (L + 4095) / 4096));
    std::vector<Part> parts((size_t)T);
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) {
        const int64_t lo = L * t / T, hi = L * (t + 1) / T;
        pool.emplace_back(parse_range, std::cref(voc), text, std::cref(line_start), lo, hi, &parts[(size_t)t]);
    }
    for (auto& th : pool) th.join();
    pylda_parsed* r = new pylda_parsed();
    int64_t D = 0, nnz = 0;
    for (const Part& p : parts) {
        D += (int64_t)p.doc_nnz.size();
        nnz += (int64_t)p.ids.size();
        r->dropped += p.dropped;
    }
    r->row_ptr.reserve((size_t)D + 1);
    r->ids.reserve((size_t)nnz);
    r->cts.reserve((size_t)nnz);
    r->row_ptr.push_back(0);
    for (const Part& p : parts) {
        for (int64_t n : p.doc_nnz) r->row_ptr.push_back(r->row_ptr.back() + n);
        r->ids.insert(r->ids.end(), p.ids.begin(), p.ids.end());
        r->cts.insert(r->cts.end(), p.cts.begin(), p.cts.end());
    }
    *out = r;
    return 0;
}

int pylda_parsed_dims(const pylda_parsed* p, int64_t* D, int64_t* nnz, int64_t* dropped) {
    if (!p) return 1;
    if (D) *D = (int64_t)p->row_ptr.size() - 1;
    if (nnz) *nnz = (int64_t)p->ids.size();
    if (dropped) *dropped = p->dropped;
    return 0;
}

int pylda_parsed_copy(const pylda_parsed* p, int64_t* row_ptr, int32_t* ids, int32_t* cts) {
    if (!p) return 1;
    if (row_ptr) memcpy(row_ptr, p->row_ptr.data(), p->row_ptr.size() * sizeof(int64_t));
    if (ids && !p->ids.empty()) memcpy(ids, p->ids.data(), p->ids.size() * sizeof(int32_t));
    if (cts && !p->cts.empty()) memcpy(cts, p->cts.data(), p->cts.size() * sizeof(int32_t));
    return 0;
}

int pylda_parsed_free(pylda_parsed* p) {
    delete p;
    return 0;
}

}  // extern "C"
