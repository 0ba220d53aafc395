"""Import the TEXT of the reference's own test document (tests/data/alps_wiki.html, the Wikipedia article "Alps",
CC BY-SA: see the reference's tests/data/ATTRIBUTION.md) as the corpus of BASELINE config 1.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden_alps.py

The HTML is reduced to visible paragraph / heading / list text with the standard library's HTMLParser (the reference
parses it with `unstructured`, which is not installed offline), whitespace-normalised and cut into pieces of at most
1000 characters at sentence boundaries.  Output: tests/golden/alps_wiki_chunks.json -- a fixture,
not code; the parity tests embed it with the CUDA encoder and compare against the fp32 oracle.
"""

from __future__ import annotations

import json
import os
import re
from html.parser import HTMLParser

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(os.environ.get("DIAL_RAG_REFERENCE_ROOT", "/root/reference"), "tests", "data", "alps_wiki.html")
OUT = os.path.join(ROOT, "tests", "golden", "alps_wiki_chunks.json")


class TextOf(HTMLParser):
    KEEP = {"p", "h1", "h2", "h3", "li"}
    SKIP = {"script", "style", "table", "sup", "nav", "footer", "figure"}

    def __init__(self):
        super().__init__()
        self.blocks, self.cur, self.depth_keep, self.depth_skip = [], [], 0, 0

    def handle_starttag(self, tag, attrs):
        if tag in self.SKIP:
            self.depth_skip += 1
        elif tag in self.KEEP:
            self.depth_keep += 1

    def handle_endtag(self, tag):
        if tag in self.SKIP and self.depth_skip:
            self.depth_skip -= 1
        elif tag in self.KEEP and self.depth_keep:
            self.depth_keep -= 1
            if self.depth_keep == 0:
                text = re.sub(r"\s+", " ", "".join(self.cur)).strip()
                if len(text) >= 40:
                    self.blocks.append(text)
                self.cur = []

    def handle_data(self, data):
        if self.depth_keep and not self.depth_skip:
            self.cur.append(data)


def pieces(block: str, limit: int = 1000):
    out, cur = [], ""
    for sentence in re.split(r"(?<=[.!?])\s+", block):
        while len(sentence) > limit:
            out.append(sentence[:limit])
            sentence = sentence[limit:]
        if cur and len(cur) + 1 + len(sentence) > limit:
            out.append(cur)
            cur = sentence
        else:
            cur = (cur + " " + sentence).strip()
    if cur:
        out.append(cur)
    return out


def main() -> None:
    with open(SRC, encoding="utf-8") as f:
        parser = TextOf()
        parser.feed(f.read())
    chunks = [p for b in parser.blocks for p in pieces(b)]
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump({"source": "epam/ai-dial-rag tests/data/alps_wiki.html (Wikipedia 'Alps', CC BY-SA 4.0)", "chunks": chunks},
                  f, ensure_ascii=False, indent=0)
    print(len(chunks), "chunks,", sum(map(len, chunks)), "characters ->", OUT)


if __name__ == "__main__":
    main()
