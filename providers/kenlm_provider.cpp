// Scorer provider for ctcdecode_b200 (include/ctcdecode_b200.h: ctcdec_scorer_hooks; ctcdecode_b200/scorer.py).
//
// The language model stays on the host behind the reference's own Scorer plugin surface: this file is the small
// C-ABI stub a maintainer of parlance/ctcdecode adds next to the UNMODIFIED scorer.cpp (+ the vendored KenLM it
// already builds) -- see INTEGRATION.md section 6.  It contains no decoder: no DecoderState, no beam search, no
// ThreadPool; the prefix beam search runs in the CUDA library.  Replaces what reference binding.cpp:143-150
// (paddle_get_scorer), :274-288 (is_character_based / get_max_order / get_dict_size / reset_params) and
// ctc_beam_search_decoder.cpp:120-137,194-206 (the two calls into the Scorer) reach.
//
// Built by providers/Makefile from the reference sources WHERE THEY LIE (nothing is copied):
//   scorer.cpp path_trie.cpp decoder_utils.cpp + third_party/kenlm  ->  providers/_build/libkenlm_provider.so
#include <cstring>
#include <string>
#include <vector>

#include "lm/config.hh"
#include "lm/model.hh"
#include "path_trie.h"
#include "scorer.h"

extern "C" {

void *ref_scorer_new(double alpha, double beta, const char *lm_path, const char *const *labels, int n_labels) {
  std::vector<std::string> vocab;
  for (int i = 0; i < n_labels; ++i) vocab.emplace_back(labels[i]);
  return new Scorer(alpha, beta, lm_path, vocab);  // reference binding.cpp:143-150
}
void ref_scorer_free(void *s) { delete static_cast<Scorer *>(s); }
int ref_scorer_is_character_based(void *s) { return static_cast<Scorer *>(s)->is_character_based(); }
size_t ref_scorer_max_order(void *s) { return static_cast<Scorer *>(s)->get_max_order(); }
size_t ref_scorer_dict_size(void *s) { return static_cast<Scorer *>(s)->get_dict_size(); }
void ref_scorer_reset_params(void *s, double a, double b) { static_cast<Scorer *>(s)->reset_params(a, b); }

// Scorer::get_log_cond_prob(Scorer::make_ngram(prefix)) for the prefix spelled by `labels`
// (reference scorer.cpp:74-93,163-194; called at ctc_beam_search_decoder.cpp:127-131)
double ref_scorer_cond_from_labels(void *scorer, const int *labels, int n) {
  Scorer *s = static_cast<Scorer *>(scorer);
  std::vector<PathTrie> chain(static_cast<size_t>(n) + 1);  // chain[0] is the root (character == -1)
  for (int k = 0; k < n; ++k) {
    chain[k + 1].character = labels[k];
    chain[k + 1].parent = &chain[k];
  }
  return s->get_log_cond_prob(s->make_ngram(&chain[n]));
}

// Scorer::get_sent_log_prob(Scorer::split_labels(prefix))  (reference scorer.cpp:95-146; called at
// ctc_beam_search_decoder.cpp:194-206)
double ref_scorer_sent_from_labels(void *scorer, const int *labels, int n) {
  Scorer *s = static_cast<Scorer *>(scorer);
  std::vector<int> v(labels, labels + n);
  return s->get_sent_log_prob(s->split_labels(v));
}

// The language model's vocabulary as KenLM enumerates it (what Scorer::load_lm collects, scorer.cpp:55-72),
// '\n'-separated into buf; returns the number of bytes needed.
size_t ref_lm_vocabulary(const char *lm_path, char *buf, size_t cap) {
  RetriveStrEnumerateVocab enumerate;
  lm::ngram::Config config;
  config.enumerate_vocab = &enumerate;
  lm::base::Model *m = lm::ngram::LoadVirtual(lm_path, config);
  delete m;
  std::string all;
  for (const std::string &w : enumerate.vocabulary) { all += w; all += '\n'; }
  if (buf && cap >= all.size() + 1) memcpy(buf, all.c_str(), all.size() + 1);
  return all.size() + 1;
}

}  // extern "C"
