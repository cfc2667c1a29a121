"""Writes tests/data/char_lm.arpa: a small order-3 CHARACTER language model (every word of its vocabulary is one
UTF-8 character, so the reference Scorer reports is_character_based(), scorer.cpp:63-71).  Deterministic: the
numbers come from a fixed integer recurrence, not from data.  "|" plays the word separator; " " is not in the model
(ARPA words cannot be blanks), so a label " " scores as out-of-vocabulary, like every label the model lacks."""
import os

CHARS = list("abcdefghijklmnop") + ["|", "'", "é"]


def main():
    state = [12345]

    def rnd():
        state[0] = (state[0] * 1103515245 + 12345) % (1 << 31)
        return state[0] / float(1 << 31)

    uni = [("<unk>", -2.5, 0.0), ("<s>", -99.0, -0.5), ("</s>", -1.6, 0.0)]
    for ch in CHARS:
        uni.append((ch, -0.9 - 1.2 * rnd(), -0.1 - 0.5 * rnd()))
    bi = {}
    for a in ["<s>"] + CHARS:
        for b in CHARS + ["</s>"]:
            if rnd() < 0.35:
                bi[(a, b)] = (-0.3 - 1.5 * rnd(), -0.05 - 0.4 * rnd())
    tri = {}
    for (a, b) in sorted(bi):
        if b == "</s>":
            continue
        for c in CHARS + ["</s>"]:
            if (b, c) in bi and rnd() < 0.3:
                tri[(a, b, c)] = -0.2 - 1.2 * rnd()
    out = ["\\data\\", "ngram 1=%d" % len(uni), "ngram 2=%d" % len(bi), "ngram 3=%d" % len(tri), "", "\\1-grams:"]
    for w, p, bo in uni:
        out.append("%.4f\t%s\t%.4f" % (p, w, bo))
    out += ["", "\\2-grams:"]
    for (a, b) in sorted(bi):
        p, bo = bi[(a, b)]
        out.append("%.4f\t%s %s\t%.4f" % (p, a, b, bo) if b != "</s>" else "%.4f\t%s %s" % (p, a, b))
    out += ["", "\\3-grams:"]
    for (a, b, c) in sorted(tri):
        out.append("%.4f\t%s %s %s" % (tri[(a, b, c)], a, b, c))
    out += ["", "\\end\\", ""]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "char_lm.arpa")
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(out))
    print(path, len(uni), len(bi), len(tri))


if __name__ == "__main__":
    main()
