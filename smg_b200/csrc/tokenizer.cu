// K1 — GPU tokenizer: vocabulary tables + pre-tokenise / BPE / compaction kernels.  See tokenizer.h and bpe.cuh.
#include "tokenizer.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_map>

#include "unicode_tables.h"

namespace smgx {

namespace {

constexpr uint32_t kInvalidTok = 0xFFFFFFFFu;

// K1a: one thread per request walks its text once: special tokens are cut out first (tiktoken `encode`: the regex only
// ever sees the text between two specials), then piece starts are flagged for the BPE kernel.
//   flags[i] = 1  piece of ordinary text starts at byte i ; 2 = first byte of a special token ; 3 = inside a special token
//   tmp_ids  = special id at a special's first byte, INVALID on its other bytes
//   totals[r] = number of special tokens of request r (the BPE kernel adds the ordinary tokens)
__global__ void __launch_bounds__(128) pretokenize_kernel(BpeView v, const uint8_t* __restrict__ text, const uint32_t* __restrict__ offsets,
                                                          uint32_t n, uint8_t* __restrict__ flags, uint32_t* __restrict__ tmp_ids,
                                                          uint32_t* __restrict__ totals) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t beg = offsets[r], end = offsets[r + 1];
    const uint8_t* s = text + beg;
    const uint32_t len = end - beg;
    uint32_t i = 0, n_special = 0;
    while (i < len) {
        // next special token at or after i
        uint32_t sp = len, sp_len = 0, sp_id = 0;
        if (v.n_special) {
            for (uint32_t k = i; k < len; ++k) {
                if (special_at(v, s, k, len, sp_len, sp_id)) { sp = k; break; }
            }
        }
        // ordinary text [i, sp): regex pieces, with sp acting as the end of the string
        while (i < sp) {
            flags[beg + i] = 1;
            i = next_piece_cl100k(s, i, sp, v.uni);
        }
        if (sp < len) {
            tmp_ids[beg + sp] = sp_id;
            flags[beg + sp] = 2;
            for (uint32_t k = 1; k < sp_len; ++k) { tmp_ids[beg + sp + k] = kInvalidTok; flags[beg + sp + k] = 3; }
            ++n_special;
            i = sp + sp_len;
        }
    }
    totals[r] = n_special;
}

// K1b: one thread per flagged byte = one piece.  Whole-piece vocabulary hit → one token; otherwise byte-pair merge in
// place (ids in tmp_ids[i..], pair ranks in tmp_rk[i..]).  Unused slots of the piece are set INVALID for the compaction.
__global__ void __launch_bounds__(256) bpe_pieces_kernel(BpeView v, const uint8_t* __restrict__ text, const uint32_t* __restrict__ offsets,
                                                         uint32_t n, uint32_t total_bytes, const uint8_t* __restrict__ flags,
                                                         uint32_t* __restrict__ tmp_ids, uint64_t* __restrict__ tmp_rk,
                                                         uint32_t* __restrict__ totals) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total_bytes; i += gridDim.x * blockDim.x) {
        if (flags[i] != 1) continue;
        // request of byte i: last r with offsets[r] <= i
        uint32_t lo = 0, hi = n;
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (offsets[m] <= i) lo = m; else hi = m; }
        const uint32_t r = lo, rend = offsets[r + 1];
        uint32_t e = i + 1;
        while (e < rend && flags[e] == 0) ++e;   // piece ends at the next piece start / special / request end
        const uint32_t plen = e - i;
        uint32_t k, id;
        if (piece_lookup(v, text + i, plen, id)) { tmp_ids[i] = id; k = 1; }
        else k = byte_pair_merge(v, text + i, plen, tmp_ids + i, tmp_rk + i);
        for (uint32_t j = i + k; j < e; ++j) tmp_ids[j] = kInvalidTok;
        atomicAdd(&totals[r], k);
    }
}

// exclusive scan of per-request token counts → token offsets (n + 1); one CTA, n ≤ a few 10^5
__global__ void __launch_bounds__(1024) scan_counts_kernel(const uint32_t* __restrict__ totals, uint32_t n, uint32_t* __restrict__ tok_offsets) {
    __shared__ uint32_t s_part[1024];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < n ? totals[i] : 0;
        s_part[threadIdx.x] = x;
        __syncthreads();
        for (uint32_t d = 1; d < 1024; d <<= 1) {
            uint32_t t = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) tok_offsets[i] = s_carry + s_part[threadIdx.x] - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry += s_part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) tok_offsets[n] = s_carry;
}

// K1c: one warp per request: stream-compact the valid token slots of its byte range into the ragged output
__global__ void __launch_bounds__(256) compact_tokens_kernel(const uint32_t* __restrict__ offsets, uint32_t n, const uint32_t* __restrict__ tmp_ids,
                                                             const uint32_t* __restrict__ tok_offsets, uint32_t* __restrict__ out_tokens) {
    const int lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t beg = offsets[r], end = offsets[r + 1];
    uint32_t out = tok_offsets[r];
    for (uint32_t base = beg; base < end; base += 32) {
        uint32_t i = base + lane;
        uint32_t v = i < end ? tmp_ids[i] : kInvalidTok;
        unsigned m = __ballot_sync(0xffffffffu, v != kInvalidTok);
        if (v != kInvalidTok) out_tokens[out + __popc(m & ((1u << lane) - 1u))] = v;
        out += __popc(m);
    }
}

uint32_t pow2_at_least(uint64_t x) { uint32_t p = 1024; while (p < x) p <<= 1; return p; }

std::string b64decode(const std::string& in) {
    static int8_t T[256];
    static bool init = false;
    if (!init) {
        memset(T, -1, sizeof(T));
        const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; ++i) T[(uint8_t)a[i]] = (int8_t)i;
        init = true;
    }
    std::string out;
    uint32_t val = 0;
    int bits = -8;
    for (unsigned char c : in) {
        if (c == '=') break;
        if (T[c] < 0) throw Error(SMGX_TOKENIZATION_ERROR, "invalid base64 in tiktoken file");
        val = (val << 6) | (uint32_t)T[c];
        bits += 6;
        if (bits >= 0) { out.push_back((char)((val >> bits) & 0xFF)); bits -= 8; }
    }
    return out;
}

}  // namespace

Tokenizer* Tokenizer::from_tiktoken_file(const std::string& path, const std::vector<std::pair<std::string, uint32_t>>& specials, bool device) {
    // load_tiktoken_bpe (tiktoken.rs:346-367): lines of `base64(token) rank`
    std::ifstream f(path);
    if (!f) throw Error(SMGX_TOKENIZATION_ERROR, "cannot open tiktoken file '" + path + "'");
    std::vector<std::string> toks;
    std::vector<uint32_t> ranks;
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        std::istringstream ls(line);
        std::string b64, rk;
        if (!(ls >> b64)) throw Error(SMGX_TOKENIZATION_ERROR, "missing token in tiktoken file");
        if (!(ls >> rk)) throw Error(SMGX_TOKENIZATION_ERROR, "missing rank in tiktoken file");
        toks.push_back(b64decode(b64));
        ranks.push_back((uint32_t)std::stoul(rk));
    }
    return new Tokenizer(toks, ranks, specials, device);
}

Tokenizer::Tokenizer(const std::vector<std::string>& tokens, const std::vector<uint32_t>& ranks,
                     const std::vector<std::pair<std::string, uint32_t>>& specials, bool device)
    : device_(device) {
    SMGX_REQUIRE(tokens.size() == ranks.size() && !tokens.empty(), "empty vocabulary");
    std::unordered_map<std::string, uint32_t> enc;
    enc.reserve(tokens.size() * 2);
    uint32_t max_len = 1, max_id = 0;
    for (size_t i = 0; i < tokens.size(); ++i) {
        SMGX_REQUIRE(!tokens[i].empty(), "empty token in vocabulary");
        enc[tokens[i]] = ranks[i];
        max_len = std::max<uint32_t>(max_len, (uint32_t)tokens[i].size());
        max_id = std::max(max_id, ranks[i]);
    }
    for (auto& sp : specials) max_id = std::max(max_id, sp.second);
    vocab_size_ = max_id + 1;

    std::vector<uint32_t> byte_token(256);
    for (int b = 0; b < 256; ++b) {
        auto it = enc.find(std::string(1, (char)b));
        if (it == enc.end()) throw Error(SMGX_TOKENIZATION_ERROR, "vocabulary lacks the single-byte token " + std::to_string(b));
        byte_token[b] = it->second;
    }
    // blob of token bytes + piece table keyed by the bytes
    std::vector<uint8_t> blob;
    std::vector<PieceSlot> pieces(pow2_at_least(tokens.size() * 2), PieceSlot{0, 0, 0, 0, 0});
    uint32_t pmask = (uint32_t)pieces.size() - 1;
    for (auto& kv : enc) {
        uint32_t off = (uint32_t)blob.size();
        blob.insert(blob.end(), kv.first.begin(), kv.first.end());
        uint64_t hsh = bytes_hash((const uint8_t*)kv.first.data(), (uint32_t)kv.first.size());
        uint32_t h = (uint32_t)(hsh >> 32) & pmask;
        while (pieces[h].hash) h = (h + 1) & pmask;
        pieces[h] = PieceSlot{hsh, kv.second, (uint32_t)kv.first.size(), off, 0};
    }
    // pair table: every split of every token into two vocab entries (rank of the concatenation = tiktoken's merge priority)
    std::vector<std::pair<uint64_t, uint32_t>> pair_list;
    for (auto& kv : enc) {
        const std::string& t = kv.first;
        for (size_t k = 1; k < t.size(); ++k) {
            auto l = enc.find(t.substr(0, k));
            if (l == enc.end()) continue;
            auto r = enc.find(t.substr(k));
            if (r == enc.end()) continue;
            pair_list.push_back({((uint64_t)l->second << 32) | r->second, kv.second});
        }
    }
    n_pairs_ = (uint32_t)pair_list.size();
    std::vector<PairSlot> pairs(pow2_at_least(pair_list.size() * 2 + 2), PairSlot{kPairEmpty, kRankMax, 0});
    uint32_t qmask = (uint32_t)pairs.size() - 1;
    for (auto& pr : pair_list) {
        uint32_t h = (uint32_t)(mix64(pr.first) >> 32) & qmask;
        while (pairs[h].key != kPairEmpty && pairs[h].key != pr.first) h = (h + 1) & qmask;
        pairs[h] = PairSlot{pr.first, pr.second, pr.second};
    }
    // specials
    std::vector<SpecialTok> sp;
    uint32_t first_bits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (auto& s : specials) {
        SMGX_REQUIRE(!s.first.empty(), "empty special token");
        uint32_t off = (uint32_t)blob.size();
        blob.insert(blob.end(), s.first.begin(), s.first.end());
        sp.push_back(SpecialTok{off, (uint32_t)s.first.size(), s.second, 0});
        uint8_t c = (uint8_t)s.first[0];
        first_bits[c >> 5] |= 1u << (c & 31);
    }
    if (!device_) return;
    auto up = [](DevBuf& d, const void* src, size_t bytes) {
        d.reserve(std::max<size_t>(bytes, 16));
        if (bytes) SMGX_CUDA(cudaMemcpy(d.ptr, src, bytes, cudaMemcpyHostToDevice));
    };
    up(d_byte_token_, byte_token.data(), 256 * 4);
    up(d_pairs_, pairs.data(), pairs.size() * sizeof(PairSlot));
    up(d_pieces_, pieces.data(), pieces.size() * sizeof(PieceSlot));
    up(d_blob_, blob.data(), blob.size());
    up(d_specials_, sp.data(), sp.size() * sizeof(SpecialTok));
    up(d_uni_lo_, kUniLo, sizeof(kUniLo));
    up(d_uni_hi_, kUniHi, sizeof(kUniHi));
    up(d_uni_cls_, kUniCls, sizeof(kUniCls));
    dview_.byte_token = d_byte_token_.as<uint32_t>();
    dview_.pairs = d_pairs_.as<PairSlot>();
    dview_.pair_mask = qmask;
    dview_.pieces = d_pieces_.as<PieceSlot>();
    dview_.piece_mask = pmask;
    dview_.blob = d_blob_.as<uint8_t>();
    dview_.max_token_len = max_len;
    dview_.specials = d_specials_.as<SpecialTok>();
    dview_.n_special = (uint32_t)sp.size();
    for (int i = 0; i < 8; ++i) dview_.special_first[i] = first_bits[i];
    dview_.uni = UnicodeView{d_uni_lo_.as<uint32_t>(), d_uni_hi_.as<uint32_t>(), d_uni_cls_.as<uint8_t>(), kUniRanges};
}

Tokenizer::~Tokenizer() {
    d_byte_token_.release(); d_pairs_.release(); d_pieces_.release(); d_blob_.release(); d_specials_.release();
    d_uni_lo_.release(); d_uni_hi_.release(); d_uni_cls_.release();
}

void Tokenizer::encode_batch(const uint8_t* d_text, const uint32_t* d_offsets, uint32_t n, uint32_t total_bytes, uint32_t* d_tokens,
                             uint32_t* d_tok_offsets, Scratch& sc, cudaStream_t stream, uint64_t* launches) const {
    if (!device_) throw Error(SMGX_DEVICE_ERROR, "tokenizer was loaded without a device (host-mirror policy): no GPU path, no CPU fallback");
    if (n == 0) return;
    sc.flags.reserve(std::max<uint32_t>(total_bytes, 1));
    sc.tmp_ids.reserve((size_t)std::max<uint32_t>(total_bytes, 1) * 4);
    sc.tmp_rk.reserve((size_t)std::max<uint32_t>(total_bytes, 1) * 8);
    sc.totals.reserve((size_t)n * 4);
    if (total_bytes) {
        SMGX_CUDA(cudaMemsetAsync(sc.flags.ptr, 0, total_bytes, stream));
    }
    pretokenize_kernel<<<(n + 127) / 128, 128, 0, stream>>>(dview_, d_text, d_offsets, n, sc.flags.as<uint8_t>(), sc.tmp_ids.as<uint32_t>(),
                                                          sc.totals.as<uint32_t>());
    SMGX_CUDA(cudaGetLastError());
    ++*launches;
    if (total_bytes) {
        unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)total_bytes + 255) / 256, 148ull * 32);
        bpe_pieces_kernel<<<grid, 256, 0, stream>>>(dview_, d_text, d_offsets, n, total_bytes, sc.flags.as<uint8_t>(), sc.tmp_ids.as<uint32_t>(),
                                                  sc.tmp_rk.as<uint64_t>(), sc.totals.as<uint32_t>());
        SMGX_CUDA(cudaGetLastError());
        ++*launches;
    }
    scan_counts_kernel<<<1, 1024, 0, stream>>>(sc.totals.as<uint32_t>(), n, d_tok_offsets);
    SMGX_CUDA(cudaGetLastError());
    ++*launches;
    compact_tokens_kernel<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, stream>>>(d_offsets, n, sc.tmp_ids.as<uint32_t>(), d_tok_offsets, d_tokens);
    SMGX_CUDA(cudaGetLastError());
    ++*launches;
}

}  // namespace smgx
