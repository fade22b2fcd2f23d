"""Text <-> token ids with the reference's semantics (src/transformer/tokenizer.cpp:235-399; host logic there and here).

encode (:247-327): optional BOS; the text is cut into UTF-8 code points (a lead byte plus up to three continuation bytes);
each is looked up in the vocabulary, where a single space is looked up as the connector piece "▁" (:236-239; there is NO
dummy-prefix space, :264-269) and an unknown code point falls back to byte tokens `byte + 3`; then the adjacent pair whose
concatenation is a vocabulary entry with the highest score is merged, leftmost first on ties, until none is left.  The
reference rescans all pairs per merge (O(n^2)); a heap keyed by (-score, position) pops the same pair every time, so
the result is identical in O(n log n).
decode (:329-343, :387-399): the piece's display text ("▁x" shown as " x" for .flm / GGUF vocabularies), minus a leading space
right after BOS (token id 1), "<0xHH>" pieces turned into the raw byte, and single non-printable bytes dropped."""
import heapq
import re
import struct

import numpy as np

CONN_TAG = "▁".encode("utf-8")
_BYTE_PIECE = re.compile(rb"<0x([0-9A-Fa-f]{1,2})")            # sscanf("<0x%02hhX>") stops caring after the hex digits
_PRINTABLE_OR_SPACE = set(range(0x20, 0x7F)) | {0x09, 0x0A, 0x0B, 0x0C, 0x0D}


class Tokenizer:
    def __init__(self, texts, scores, show=None, conn_tag=CONN_TAG, bos=1, eos=2, pad=0):
        """texts: the pieces (bytes or str) indexed by token id; show: their display form (default: the same)"""
        enc = lambda t: t.encode("utf-8", "surrogateescape") if isinstance(t, str) else bytes(t)      # noqa: E731
        tb = [enc(t) for t in texts]
        self.texts = tb
        self.show = tb if show is None else [enc(t) for t in show]
        self.scores = np.asarray(scores, np.float32)
        self.bos, self.eos, self.pad = bos, eos, pad
        self.text2id = {t: i for i, t in enumerate(tb)}                # a repeated piece keeps its LAST id (:153-155)
        conn = conn_tag.encode("utf-8") if isinstance(conn_tag, str) else conn_tag
        self.underline_id = self.text2id.get(conn, -1)

    # ---- constructors for the three vocabulary sources -----------------------------------------------------------------
    @classmethod
    def from_flm_vocab(cls, vocab):
        """vocab dict returned by flm.read_flm (Tokenizer::set, tokenizer.cpp:57-72): unset special ids are -1"""
        sp = vocab.get("special", {})
        return cls(vocab["texts"], vocab["scores"], vocab.get("show"), vocab.get("conn_tag", CONN_TAG),
                   sp.get("bos", -1), sp.get("eos", -1), sp.get("pad", -1))

    @classmethod
    def from_gguf_vocab(cls, vocab):
        """vocab dict returned by gguf_file.read_gguf (set_token_texts, tokenizer.cpp:74-120): connector pieces are displayed
        with a leading space; special ids default to 1 / 2 / 0"""
        texts = [t.encode("utf-8", "surrogateescape") for t in vocab["texts"]]
        show = [b" " + t[len(CONN_TAG):] if t.startswith(CONN_TAG) else t for t in texts]
        sp = vocab.get("special", {})
        return cls(texts, vocab["scores"], show, CONN_TAG, sp.get("bos", 1), sp.get("eos", 2), sp.get("pad", 0))

    @classmethod
    def from_tokenizer_bin(cls, path, vocab_size):
        """llama2.c tokenizer.bin (Tokenizer::load, :159-233): i32 max_token_length, then {f32 score, i32 len, bytes} per
        token.  No connector tag is set on this path, so a space is NOT mapped to "▁" (it falls back to its byte token)."""
        texts, scores = [], []
        with open(path, "rb") as f:
            f.read(4)
            for _ in range(vocab_size):
                score, n = struct.unpack("<fi", f.read(8))
                t = f.read(n)
                if len(t) != n:
                    raise ValueError(f"{path}: truncated tokenizer file")
                scores.append(score)
                texts.append(t.split(b"\0", 1)[0])                     # C strings end at the first NUL
        return cls(texts, scores, None, b"", 1, 2, 0)

    # ---- encode ---------------------------------------------------------------------------------------------------------
    def _search(self, piece):
        if piece == b" ":
            return self.underline_id
        return self.text2id.get(piece, -1)

    def encode(self, text, add_bos=True, add_eos=False):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        b = b.split(b"\0", 1)[0]                                          # the reference is handed a C string
        if not b:
            return []
        toks = [self.bos] if add_bos else []
        cp = bytearray()
        for i, c in enumerate(b):
            if (c & 0xC0) != 0x80:
                cp.clear()
            cp.append(c)
            nxt = b[i + 1] if i + 1 < len(b) else 0
            if (nxt & 0xC0) == 0x80 and len(cp) < 4:
                continue
            tid = self._search(bytes(cp))
            if tid >= 0:
                toks.append(tid)
            else:
                toks.extend(x + 3 for x in cp)
            cp.clear()
        toks = self._merge(toks)
        if add_eos:
            toks.append(self.eos)
        return toks

    def _merge(self, toks):
        n = len(toks)
        if n < 2:
            return toks
        sym = list(toks)
        prev = list(range(-1, n - 1))
        nxt = list(range(1, n + 1))
        nxt[-1] = -1
        alive = [True] * n
        heap = []

        def push(i):
            j = nxt[i]
            if j < 0 or sym[i] >= len(self.texts) or sym[j] >= len(self.texts):      # byte fallback ids beyond a tiny vocabulary
                return
            tid = self._search(self.texts[sym[i]] + self.texts[sym[j]])
            if tid != -1 and self.scores[tid] > np.float32(-1e10):
                heapq.heappush(heap, (-float(self.scores[tid]), i, sym[i], sym[j], tid))

        for i in range(n - 1):
            push(i)
        while heap:
            _, i, a, b, tid = heapq.heappop(heap)
            j = nxt[i] if alive[i] else -1
            if j < 0 or sym[i] != a or sym[j] != b:
                continue                                                # a stale entry: one of the two has been merged away
            sym[i] = tid
            alive[j] = False
            nxt[i] = nxt[j]
            if nxt[j] >= 0:
                prev[nxt[j]] = i
            if prev[i] >= 0:
                push(prev[i])
            push(i)
        out, i = [], 0
        while i >= 0:
            out.append(sym[i])
            i = nxt[i]
        return out

    # ---- decode ---------------------------------------------------------------------------------------------------------
    def decode_piece(self, token, prev_token=-1):
        if token < 0 or token >= len(self.show):
            return b""
        piece = self.show[token]
        if prev_token == 1 and piece[:1] == b" ":
            piece = piece[1:]
        m = _BYTE_PIECE.match(piece)
        if m:
            piece = bytes([int(m.group(1), 16)])
            if piece == b"\0":
                return b""                                              # an empty C string
        if not piece:
            return b""
        if len(piece) == 1 and piece[0] not in _PRINTABLE_OR_SPACE:
            return b""
        return piece

    def decode(self, tokens):
        out, prev = [], -1
        for t in tokens:
            out.append(self.decode_piece(int(t), prev))
            prev = int(t)
        return b"".join(out)
