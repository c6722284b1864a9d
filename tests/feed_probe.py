"""Helper for tests/test_feed_cpu.py — run as a SUBPROCESS (it registers import stubs), only where the reference tree exists
(the build container; never on the GPU box).  Differential check on fresh random cases, beyond the committed fixture: the
reference's TextMaskingGenerator / preprocess / get_image_attns / RandomAugment sampling and xfm_b200.feed, same seeds.
Prints one JSON object."""
import contextlib
import io
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import make_golden_feed as G
    from feed_stub import StubTokenizer, WORDS
    from xfm_b200 import feed

    pd = G.reference_module()
    import dataset.randaugment as ra
    quiet = contextlib.redirect_stdout(io.StringIO())
    gen = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 99)
    res = dict(masker=0, preprocess=0, image_atts=0, randaugment=0, mismatches=[])
    for style in ("roberta", "bert"):
        tok = StubTokenizer(style)
        for _ in range(12):
            cfg = (gen.choice([0.15, 0.25, 0.4, 0.6]), gen.randint(1, 10), gen.choice([0.0, 0.2, 0.7, 1.0]), gen.randint(1, 4),
                   gen.random() < 0.5, style == "roberta" and gen.random() < 0.5)
            with quiet:
                ref = pd.TextMaskingGenerator(tok, *cfg[:5], use_roberta=cfg[5])
            mine = feed.TextMasker(tok, *cfg[:5], use_roberta=cfg[5])
            ds = G.bare(pd.ImageTextJsonDataset, tokenized=False, language_chosen=None, max_words=gen.randint(3, 12),
                        max_tokens=gen.randint(4, 16), max_masks=cfg[1], tokenizer=tok, cls_token=tok.cls_token,
                        eos_token=tok.sep_token, pad_token_id=tok.pad_token_id, add_eos=True, mask_generator=ref, PAD_mask=-100)
            tp = feed.TextPreprocessor(tok, mine, max_tokens=ds.max_tokens, max_masks=cfg[1], max_words=ds.max_words)
            for _ in range(25):
                text = " ".join(gen.choice(WORDS) for _ in range(gen.randint(1, 14)))
                tokens = [tok.cls_token] + tok.tokenize(text) + [tok.sep_token]
                seed = gen.randrange(1 << 30)
                random.seed(seed)
                a = ref(list(tokens)), random.random()
                random.seed(seed)
                b = mine(list(tokens)), random.random()
                res["masker"] += 1
                if a != b:
                    res["mismatches"].append(["masker", style, cfg, seed, tokens])
                random.seed(seed)
                a = [list(map(int, r)) for r in ds.preprocess(text)], random.random()
                random.seed(seed)
                b = [list(r) for r in tp.preprocess(text)], random.random()
                res["preprocess"] += 1
                if a != b:
                    res["mismatches"].append(["preprocess", style, cfg, seed, text])
    for _ in range(400):
        ps = gen.choice([14, 16, 32])
        n = gen.choice([7, 14, 24])
        r = ps * n
        x, y = gen.uniform(0, r), gen.uniform(0, r)
        w, h = gen.uniform(1e-3, r - x + 2), gen.uniform(1e-3, r - y + 2)
        ds = G.bare(pd.RegionTextJsonDataset, patch_size=ps, num_patch=n)
        res["image_atts"] += 1
        if ds.get_image_attns(x, y, w, h) != feed.region_image_atts(x, y, w, h, ps, n):
            res["mismatches"].append(["image_atts", ps, n, x, y, w, h])
    fired = []
    for name in list(ra.func_dict):
        ra.func_dict[name] = (lambda name: lambda img, *args: (fired.append((name, args)), img)[1])(name)
    for _ in range(40):
        N, M = gen.randint(1, 4), gen.randint(0, 10)
        augs = gen.sample(list(feed.RandAugmentSampler.ALL), gen.randint(1, 14)) if gen.random() < 0.7 else []
        seed = gen.randrange(1 << 30)
        aug, mine = ra.RandomAugment(N, M, isPIL=False, augs=augs), feed.RandAugmentSampler(N, M, augs)
        for _ in range(10):
            np.random.seed(seed)
            del fired[:]
            aug(np.zeros((2, 2, 3), np.uint8))
            a = list(fired), float(np.random.random())
            np.random.seed(seed)
            b = mine.sample(), float(np.random.random())
            seed += 1
            res["randaugment"] += 1
            if a != b:
                res["mismatches"].append(["randaugment", N, M, augs, seed - 1])
    res["mismatches"] = res["mismatches"][:5]
    print("PROBE_JSON " + json.dumps(res))


if __name__ == "__main__":
    main()
