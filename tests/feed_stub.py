"""A deterministic stand-in for the HuggingFace tokenizers the loader restatement is tested with (vocabulary files are not in
the image): whitespace words cut into pieces of at most 3 characters, byte-level-BPE style ('Ġ' starts a word) or WordPiece
style ('##' continues one).  Used by tools/make_golden_feed.py (which drives the REFERENCE code with it) and by
tests/test_feed_cpu.py (which drives xfm_b200.feed with it)."""

WORDS = ("a an the of on in at left right man woman dog cat horse table plate riding holding standing sitting red blue green "
         "yellow small large wooden street picture image contains see we can person skateboard umbrella kitchen window "
         "mountain river beautiful photograph zebra giraffe elephant sandwich broccoli refrigerator").split()


class StubTokenizer:
    def __init__(self, style="roberta"):
        self.style = style
        if style == "roberta":
            self.cls_token, self.sep_token, self.pad_token, self.mask_token, unk = "<s>", "</s>", "<pad>", "<mask>", "<unk>"
            specials = [self.cls_token, self.pad_token, self.sep_token, unk]
        else:
            self.cls_token, self.sep_token, self.pad_token, self.mask_token, unk = "[CLS]", "[SEP]", "[PAD]", "[MASK]", "[UNK]"
            specials = [self.pad_token, unk, self.cls_token, self.sep_token]
        self.unk_token = unk
        pieces = []
        for w in WORDS:
            for p in self._pieces(w):
                if p not in pieces:
                    pieces.append(p)
        tokens = specials + pieces + [self.mask_token]
        self.vocab = {t: i for i, t in enumerate(tokens)}
        self.pad_token_id, self.mask_token_id = self.vocab[self.pad_token], self.vocab[self.mask_token]
        self.bos_token, self.eos_token = self.cls_token, self.sep_token

    def _pieces(self, word):
        cut = [word[i:i + 3] for i in range(0, len(word), 3)]
        if self.style == "roberta":
            return ["Ġ" + cut[0]] + cut[1:]
        return [cut[0]] + ["##" + c for c in cut[1:]]

    def get_vocab(self):
        return dict(self.vocab)

    def tokenize(self, text):
        out = []
        for w in text.split():
            out += self._pieces(w)
        return out

    def convert_tokens_to_ids(self, tokens):
        unk = self.vocab[self.unk_token]
        return [self.vocab.get(t, unk) for t in tokens]
