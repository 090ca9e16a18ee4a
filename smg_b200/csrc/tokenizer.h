// K1 — GPU tokenizer (tiktoken-style BPE): vocabulary tables on the device + the batch tokenize launch.
// Host-side counterpart of TiktokenTokenizer (crates/tokenizer/src/tiktoken.rs:132-462): load_tiktoken_bpe (:346-367),
// the special-token encoder (:234-238) and encode → CoreBPE::encode_with_special_tokens (:444-462).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "bpe.cuh"
#include "common.h"

namespace smgx {

class Tokenizer {
public:
    // tokens[i] = bytes of the vocab entry with rank/id ranks[i]; specials = (string, id)
    Tokenizer(const std::vector<std::string>& tokens, const std::vector<uint32_t>& ranks,
              const std::vector<std::pair<std::string, uint32_t>>& specials, bool device);
    // HuggingFace `tokenizers` BPE (byte-level): merge priority = position in `merges` (pairs of ids), only listed pairs merge;
    // ignore_merges = look the whole piece up in the vocabulary first (models::bpe ignore_merges, Llama 3).
    Tokenizer(const std::vector<std::string>& tokens, const std::vector<uint32_t>& ids, const std::vector<std::pair<uint32_t, uint32_t>>& merges,
              bool ignore_merges, const std::vector<std::pair<std::string, uint32_t>>& specials, bool device);
    ~Tokenizer();
    static Tokenizer* from_tiktoken_file(const std::string& path, const std::vector<std::pair<std::string, uint32_t>>& specials, bool device);

    BpeView view() const { return dview_; }
    uint32_t vocab_size() const { return vocab_size_; }
    uint32_t n_pairs() const { return n_pairs_; }

    // Scratch needed for a batch of `total_bytes` of text and `n` requests.
    struct Scratch { DevBuf flags, tmp_ids, tmp_rk, totals, pieces, n_pieces, miss; };

    // Enqueue pre-tokenise + BPE + compaction on `stream`.
    //   d_text/d_offsets (n+1): ragged UTF-8; d_tokens: capacity ≥ total_bytes u32; d_tok_offsets: n+1 u32 (written).
    // Text bytes live at d_text[first_byte .. total_bytes) (absolute offsets as given in d_offsets).
    // `max_len` = longest request in bytes (sizes the 2-D grids).
    void encode_batch(const uint8_t* d_text, const uint32_t* d_offsets, uint32_t n, uint32_t first_byte, uint32_t total_bytes, uint32_t max_len,
                      uint32_t* d_tokens, uint32_t* d_tok_offsets, Scratch& sc, cudaStream_t stream, uint64_t* launches) const;

private:
    void build(const std::vector<std::string>& tokens, const std::vector<uint32_t>& ranks, const std::vector<std::pair<uint32_t, uint32_t>>* merges,
               bool whole_piece, const std::vector<std::pair<std::string, uint32_t>>& specials);
    BpeView dview_{};
    uint32_t vocab_size_ = 0, n_pairs_ = 0;
    bool device_ = false;
    DevBuf d_byte_token_, d_pairs_, d_pieces_, d_blob_, d_specials_, d_uni_lo_, d_uni_hi_, d_uni_cls_;
};

}  // namespace smgx
