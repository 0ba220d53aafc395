"""A small self-written text corpus + a deterministic WordPiece vocabulary for the boundary
tests (the real bge-small-en vocab and the reference's tests/data are not available on the
GPU box).  Special tokens sit at BERT's ids: [PAD]=0 [UNK]=100 [CLS]=101 [SEP]=102 [MASK]=103."""

from __future__ import annotations

import os
import re

_TOPICS = {
    "climate": "The climate of the high mountains changes quickly with altitude. Valleys stay mild while "
               "the summits keep snow and ice through the summer, and strong winds bring sudden storms.",
    "glacier": "Glaciers carve wide valleys as they move. Meltwater from the ice feeds cold rivers and "
               "lakes, and the retreat of the glaciers is measured every year by surveyors.",
    "flora": "Alpine meadows carry gentian, edelweiss and dwarf pine. Above the tree line only mosses, "
             "lichens and a few hardy flowers survive the frost and the short growing season.",
    "fauna": "Ibex and chamois climb the steep rock faces, marmots whistle from their burrows and golden "
             "eagles circle above the ridges looking for prey.",
    "tourism": "Ski resorts, mountain railways and hiking huts bring millions of visitors. Tourism is the "
               "largest source of income for many small villages in the range.",
    "history": "Traders and armies crossed the passes for thousands of years. Roman roads, medieval mule "
               "tracks and modern tunnels follow almost the same lines through the mountains.",
    "geology": "The range was folded when two continental plates collided. Layers of limestone, granite "
               "and gneiss were pushed over one another and lifted far above the old sea floor.",
    "rivers": "Four large rivers rise near the central peaks and flow to four different seas. Dams and "
              "hydroelectric plants use the steep gradient to produce electricity.",
    "farming": "Farmers move their cattle to the high pastures in summer and make hard cheese in small "
               "dairies. Hay is cut on slopes too steep for machines.",
    "language": "German, French, Italian, Slovene and Romansh are spoken in neighbouring valleys, and "
                "dialects can change from one village to the next.",
}

_DETAILS = [
    "Measurements go back more than a century.", "Local guides tell the story differently.",
    "The pattern is strongest on the northern side.", "Researchers still debate the cause.",
    "Old maps show a very different picture.", "The effect is easiest to see in late spring.",
    "Newer studies confirm the early reports.", "Few visitors ever notice it.",
]

CHUNKS = []
for _i in range(6):
    for _topic, _text in _TOPICS.items():
        CHUNKS.append(f"{_topic.title()} note {_i + 1}.\n\n{_text} {_DETAILS[(_i * 3 + len(_topic)) % len(_DETAILS)]}")

QUERIES = [
    "what is the climate in the mountains?",
    "which animals live on the steep rock faces?",
    "how do farmers make cheese in summer?",
    "why are the glaciers retreating?",
]


def alps_wiki_chunks():
    """BASELINE config 1 corpus: the text of the reference's own tests/data/alps_wiki.html in pieces of <= 1000
    characters (fixture made by oracle/make_golden_alps.py; the real bge vocabulary is not available offline, so the
    tests tokenise it with a deterministic WordPiece vocabulary built from the text itself)."""
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "alps_wiki_chunks.json"), encoding="utf-8") as f:
        return json.load(f)["chunks"]


# the reference's own retriever test asks the first of these (tests/test_retrievers.py:90-92)
ALPS_QUERIES = [
    "what is the climate in the alps?",
    "which is the highest mountain of the alps?",
    "how were the alps formed?",
    "which animals live in the alps?",
]


def build_vocab_file(directory: str, texts=None) -> str:
    words = set()
    corpus = CHUNKS + QUERIES if texts is None else list(texts)
    for text in corpus + ["Represent this question for searching relevant passages: hello world a b"]:
        words.update(re.findall(r"[a-z]+|[^a-z\s]", text.lower()))
    vocab = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    letters = "abcdefghijklmnopqrstuvwxyz"
    vocab += list(letters) + ["##" + c for c in letters] + list("0123456789") + ["##" + c for c in "0123456789"] + list(".,:;?!'\"-()")
    # only part of the words become whole tokens so that real word-piece splitting happens
    for w in sorted(words):
        if len(w) > 1 and (len(w) <= 5 or sum(map(ord, w)) % 3):
            vocab.append(w)
    seen, uniq = set(), []
    for t in vocab:
        if t not in seen:
            seen.add(t)
            uniq.append(t)
    path = os.path.join(directory, "vocab.txt")
    with open(path, "w") as f:
        f.write("\n".join(uniq) + "\n")
    return path
