"""Mirror of tokenizer.mojo: id -> text through vocab.txt (host-side string work, no GPU)."""
from __future__ import annotations

from typing import Iterable, List


class Tokenizer:
    def __init__(self, path: str):
        with open(path, "r", encoding="utf-8") as f:
            self.vocab: List[str] = f.read().split("\n")  # tokenizer.mojo:11-13

    def decode(self, tokens: Iterable[int]) -> str:
        """tokenizer.mojo:15-28: drop <|...|> specials, 'Ġ' -> space, literal \\n -> newline."""
        out = []
        for t in tokens:
            if 0 <= t < len(self.vocab):
                tok = self.vocab[t]
                if not (tok.startswith("<|") and tok.endswith("|>")):
                    out.append(tok.replace("Ġ", " ").replace("\\n", "\n"))
        return "".join(out)
