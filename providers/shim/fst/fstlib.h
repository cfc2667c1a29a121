// TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product.
//
// Minimal stand-in for the OpenFST 1.6.7 surface that parlance/ctcdecode touches
// (path_trie.h:10,36-38,63-67; path_trie.cpp:59-96,165-174; decoder_utils.cpp:147-162;
//  scorer.cpp:196-230; ctc_beam_search_decoder.cpp:12,15,46-52).  OpenFST is not vendored
// by the reference (setup.py:26-29 downloads it) and there is no network here.
//
// The reference uses the FST only as a *dictionary acceptor*: words are added as linear
// chains from the start state, then RmEpsilon -> Determinize -> Minimize, and the decoder
// asks two questions: "is there an arc `label` out of state s?" and "is the target state
// final?".  Both answers depend on the accepted language only, so a deterministic trie
// (what Determinize yields for a union of chains, before state merging) is observably
// identical; state ids are never exposed.
#ifndef ORACLE_SHIM_FST_FSTLIB_H_
#define ORACLE_SHIM_FST_FSTLIB_H_

#include <algorithm>
#include <cmath>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace fst {

class TropicalWeight {
 public:
  TropicalWeight() : v_(std::numeric_limits<float>::infinity()) {}
  TropicalWeight(float v) : v_(v) {}  // NOLINT (implicit: reference passes literal 0)
  static TropicalWeight Zero() { return TropicalWeight(std::numeric_limits<float>::infinity()); }
  static TropicalWeight One() { return TropicalWeight(0.0f); }
  bool operator!=(const TropicalWeight &o) const { return v_ != o.v_; }
  bool operator==(const TropicalWeight &o) const { return v_ == o.v_; }
 private:
  float v_;
};

struct StdArc {
  typedef TropicalWeight Weight;
  typedef int StateId;
  typedef int Label;
  StdArc() : ilabel(0), olabel(0), weight(), nextstate(-1) {}
  StdArc(Label i, Label o, Weight w, StateId n) : ilabel(i), olabel(o), weight(w), nextstate(n) {}
  Label ilabel;
  Label olabel;
  Weight weight;
  StateId nextstate;
};

class StdVectorFst {
 public:
  typedef int StateId;
  typedef StdArc Arc;
  StdVectorFst() : start_(-1) {}
  StdVectorFst *Copy(bool = false) const { return new StdVectorFst(*this); }
  StateId Start() const { return start_; }
  void SetStart(StateId s) { start_ = s; }
  StateId NumStates() const { return static_cast<StateId>(arcs_.size()); }
  StateId AddState() {
    arcs_.emplace_back();
    final_.push_back(TropicalWeight::Zero());
    return static_cast<StateId>(arcs_.size()) - 1;
  }
  void AddArc(StateId s, const StdArc &a) { arcs_[s].push_back(a); }
  void SetFinal(StateId s, TropicalWeight w) { final_[s] = w; }
  TropicalWeight Final(StateId s) const { return final_[s]; }
  const std::vector<StdArc> &Arcs(StateId s) const { return arcs_[s]; }
 private:
  StateId start_;
  std::vector<std::vector<StdArc>> arcs_;
  std::vector<TropicalWeight> final_;
};

enum MatchType { MATCH_INPUT = 1, MATCH_OUTPUT = 2 };

template <class F>
class SortedMatcher {
 public:
  SortedMatcher(const F &fst, MatchType) : fst_(fst), state_(0), hit_(nullptr) {}
  void SetState(typename F::StateId s) { state_ = s; }
  bool Find(int label) {
    for (const auto &a : fst_.Arcs(state_)) {
      if (a.ilabel == label) {
        hit_ = &a;
        return true;
      }
    }
    hit_ = nullptr;
    return false;
  }
  const StdArc &Value() const { return *hit_; }
 private:
  const F &fst_;
  typename F::StateId state_;
  const StdArc *hit_;
};

inline void RmEpsilon(StdVectorFst *) {}  // the chains contain no epsilon arcs

// Subset construction specialised to "union of chains from one start state": merge arcs with
// equal labels recursively (= build the prefix trie); a merged state is final if any member is.
inline void Determinize(const StdVectorFst &in, StdVectorFst *out) {
  *out = StdVectorFst();
  if (in.NumStates() == 0) return;
  std::vector<std::pair<std::vector<int>, int>> work;  // (subset of input states, output state)
  int s0 = out->AddState();
  out->SetStart(s0);
  work.push_back({{in.Start()}, s0});
  while (!work.empty()) {
    auto item = work.back();
    work.pop_back();
    std::map<int, std::vector<int>> by_label;
    for (int s : item.first) {
      if (in.Final(s) != TropicalWeight::Zero()) out->SetFinal(item.second, TropicalWeight::One());
      for (const auto &a : in.Arcs(s)) by_label[a.ilabel].push_back(a.nextstate);
    }
    for (auto &kv : by_label) {
      int d = out->AddState();
      out->AddArc(item.second, StdArc(kv.first, kv.first, TropicalWeight::One(), d));
      work.push_back({kv.second, d});
    }
  }
}

inline void Minimize(StdVectorFst *) {}  // state merging is unobservable through the matcher

}  // namespace fst

#endif  // ORACLE_SHIM_FST_FSTLIB_H_
