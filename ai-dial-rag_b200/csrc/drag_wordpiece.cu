// Host-side WordPiece tokenisation for the encoder's input (at ~57k chunks/s the GPU eats
// ~15 M tokens/s; the Python-level tokenizer call is the next bottleneck of the text -> embedding path).
//
// This is the ASCII fast path of the uncased BERT tokenizer the reference uses inside sentence-transformers
// (transformers BertTokenizerFast == HF `tokenizers` BertNormalizer + BertPreTokenizer + WordPiece + the
// [CLS] ... [SEP] template, truncation to max_length):
//   clean text   control characters are dropped, \t \n \r become spaces
//   lower-case   A-Z -> a-z (nothing to strip: ASCII has no accents)
//   pre-tokenise split on spaces; every ASCII punctuation character (33-47, 58-64, 91-96, 123-126) is a word
//   WordPiece    greedy longest-match-first with the "##" continuation prefix; a word of more than 100
//                characters, or one with an unmatchable remainder, becomes [UNK]
//   template     [CLS] ids[: max_length - 2] [SEP]
// A text with any non-ASCII byte (Unicode normalisation, CJK spacing, ...) or a literal special token
// ("[SEP]" etc. are matched verbatim by the reference tokenizer) is FLAGGED and left to the caller, who runs
// the reference tokenizer on it -- the results are identical by construction, the common case is fast.
// Pure host code (std::thread over the texts); lives in the same C-ABI library as the CUDA kernels.
#include <stdint.h>
#include <string.h>

#include <fstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "drag_common.cuh"

struct drag_wordpiece {
  std::unordered_map<std::string, int32_t> vocab;
  int32_t unk = -1, cls = -1, sep = -1;
  int max_word_chars = 100;
  size_t longest_piece = 1;
  bool lowercase = true;
};

namespace {

inline bool is_punct(unsigned char c) { return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126); }
inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
inline bool is_control(unsigned char c) { return (c < 32 && !is_space(c)) || c == 127; }

// ids of one word appended to out; false if the word maps to [UNK]
bool wordpiece(const drag_wordpiece* tk, const char* w, size_t n, std::vector<int32_t>& out, std::string& scratch) {
  if ((int)n > tk->max_word_chars) return false;
  const size_t mark = out.size();
  size_t start = 0;
  while (start < n) {
    size_t end = n;
    if (end - start > tk->longest_piece) end = start + tk->longest_piece;
    int32_t id = -1;
    for (; end > start; --end) {
      scratch.clear();
      if (start > 0) scratch.append("##");
      scratch.append(w + start, end - start);
      auto it = tk->vocab.find(scratch);
      if (it != tk->vocab.end()) { id = it->second; break; }
    }
    if (id < 0) { out.resize(mark); return false; }
    out.push_back(id);
    start = end;
  }
  return true;
}

// returns the number of ids written (<= max_len), or -1 when the text has to go through the reference tokenizer
int encode_one(const drag_wordpiece* tk, const char* text, size_t n, int max_len, int32_t* out, std::vector<int32_t>& ids,
               std::string& word, std::string& scratch) {
  ids.clear();
  for (size_t i = 0; i < n; ++i) {
    const unsigned char c = (unsigned char)text[i];
    if (c >= 128) return -1;
    if (c == '[') {   // literal special tokens are matched verbatim by the reference tokenizer
      static const char* specials[] = {"[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"};
      for (const char* s : specials) {
        const size_t ls = strlen(s);
        if (n - i >= ls && memcmp(text + i, s, ls) == 0) return -1;
      }
    }
  }
  const int budget = max_len - 2;
  auto flush = [&]() {
    if (word.empty()) return;
    if ((int)ids.size() < budget && !wordpiece(tk, word.data(), word.size(), ids, scratch)) ids.push_back(tk->unk);
    word.clear();
  };
  word.clear();
  for (size_t i = 0; i < n && (int)ids.size() < budget; ++i) {
    unsigned char c = (unsigned char)text[i];
    if (is_control(c)) continue;
    if (is_space(c)) { flush(); continue; }
    if (is_punct(c)) {
      flush();
      if ((int)ids.size() < budget) {
        word.assign(1, (char)c);
        flush();
      }
      continue;
    }
    if (tk->lowercase && c >= 'A' && c <= 'Z') c = (unsigned char)(c + 32);
    word.push_back((char)c);
  }
  flush();
  int m = (int)ids.size() < budget ? (int)ids.size() : budget;
  if (m < 0) m = 0;
  out[0] = tk->cls;
  memcpy(out + 1, ids.data(), (size_t)m * 4);
  out[m + 1] = tk->sep;
  return m + 2;
}

}  // namespace

// vocab_path: BERT vocab.txt (one token per line, id = line number).  Reference: the checkpoint's tokenizer files
// read by SentenceTransformer(...) in aidial_rag/embeddings/embeddings.py:57-65.
extern "C" int drag_wordpiece_create(const char* vocab_path, int lowercase, drag_wordpiece** out) {
  DRAG_REQUIRE(vocab_path && out, "drag_wordpiece_create: null pointer");
  *out = nullptr;
  std::ifstream f(vocab_path);
  if (!f) return drag::fail(DRAG_ERR_INVALID, "drag_wordpiece_create: cannot open %s", vocab_path);
  drag_wordpiece* tk = new (std::nothrow) drag_wordpiece();
  if (!tk) return drag::fail(DRAG_ERR_NOMEM, "drag_wordpiece_create: out of host memory");
  tk->lowercase = lowercase != 0;
  std::string line;
  int32_t id = 0;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    tk->vocab.emplace(line, id);
    const size_t piece = line.rfind("##", 0) == 0 ? line.size() - 2 : line.size();
    if (piece > tk->longest_piece) tk->longest_piece = piece;
    ++id;
  }
  auto find = [&](const char* s) { auto it = tk->vocab.find(s); return it == tk->vocab.end() ? -1 : it->second; };
  tk->unk = find("[UNK]"); tk->cls = find("[CLS]"); tk->sep = find("[SEP]");
  if (tk->unk < 0 || tk->cls < 0 || tk->sep < 0) {
    delete tk;
    return drag::fail(DRAG_ERR_INVALID, "drag_wordpiece_create: %s lacks [UNK] / [CLS] / [SEP]", vocab_path);
  }
  *out = tk;
  return DRAG_OK;
}

extern "C" int drag_wordpiece_destroy(drag_wordpiece* tk) {
  delete tk;
  return DRAG_OK;
}

// texts: n_texts UTF-8 strings concatenated in `bytes`, text i = bytes[offsets[i] : offsets[i+1]].
// out_ids: int32 [n_texts][max_len] (row i holds out_len[i] ids: [CLS] ... [SEP]); out_len[i] = -1 flags a text
// the caller must run through the reference tokenizer (non-ASCII bytes or a literal special token).
extern "C" int drag_wordpiece_encode(const drag_wordpiece* tk, const char* bytes, const int64_t* offsets, int n_texts,
                                     int max_len, int n_threads, int32_t* out_ids, int32_t* out_len) {
  DRAG_REQUIRE(tk && offsets && out_ids && out_len && n_texts >= 0, "drag_wordpiece_encode: null pointer");
  DRAG_REQUIRE(max_len >= 2, "drag_wordpiece_encode: max_len must be >= 2 ([CLS] [SEP])");
  if (n_texts == 0) return DRAG_OK;
  DRAG_REQUIRE(bytes || offsets[n_texts] == offsets[0], "drag_wordpiece_encode: null text bytes");
  int workers = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (workers < 1) workers = 1;
  if (workers > n_texts) workers = n_texts;
  auto run = [&](int w) {
    std::vector<int32_t> ids;
    std::string word, scratch;
    ids.reserve((size_t)max_len);
    // interleaved assignment: neighbouring texts (similar lengths after sorting) spread over the workers
    for (int i = w; i < n_texts; i += workers)
      out_len[i] = encode_one(tk, bytes + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), max_len, out_ids + (size_t)i * max_len,
                              ids, word, scratch);
  };
  if (workers == 1) {
    run(0);
  } else {
    std::vector<std::thread> pool;
    pool.reserve((size_t)workers);
    for (int w = 0; w < workers; ++w) pool.emplace_back(run, w);
    for (auto& t : pool) t.join();
  }
  return DRAG_OK;
}
