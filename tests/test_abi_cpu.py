"""CPU-side checks of the drop-in boundary: the C-ABI shared library loads without a GPU and exports every entry point
include/xfm_b200.h declares; compute calls fail loudly (no CPU fallback); host-side helpers of the module API."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "xfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xfm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from xfm_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    handle = ctypes.CDLL(lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 40, names
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    assert lib.load().xfm_version() >= 1


def test_no_cpu_fallback():
    """Without a CUDA device every op raises instead of silently computing somewhere else."""
    from xfm_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("needs a GPU-less host")
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        lib.lib()
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        lib.gemm(a, a)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "xfm_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read().replace("the oracle", ""), f


def test_closed_form_relative_position_index_matches_beit():
    """attention_tc.cu evaluates beit2.py:94-116's index arithmetically; the closed form must equal the constructed buffer."""
    from oracle.xfm_oracle import relative_position_index
    from xfm_b200.encoders import closed_form_rel_index
    for ws in (4, 7, 12, 14, 24):
        assert torch.equal(closed_form_rel_index(ws), relative_position_index(ws))


def test_csr_inverse_groups_samples_by_image():
    from xfm_b200.encoders import csr_inverse
    kv = torch.tensor([0, 2, 1, 0, 2, 2, 0], dtype=torch.int32)
    off, smp = csr_inverse(kv, 4)
    assert off.tolist() == [0, 3, 4, 7, 7]
    assert smp.tolist() == [0, 3, 6, 2, 1, 4, 5]


def test_masking_sampler_consumes_rng_like_the_reference(golden_dir):
    import random

    import numpy as np

    from xfm_b200.masking import BlockMaskSampler
    g = torch.load(os.path.join(golden_dir, "masks.pt"), weights_only=False)
    (size, n, mn, seed), want = next(iter(g.items()))
    random.seed(seed)
    np.random.seed(seed)
    s = BlockMaskSampler(size, n, mn)
    got = np.stack([s() for _ in range(want.shape[0])])
    assert np.array_equal(got, want.numpy())


def test_itm_eval_metrics():
    """Retrieval.py:186-231 recall metrics against a literal restatement of the reference's loops."""
    import numpy as np
    from xfm_b200.retrieval_eval import itm_eval
    rng = np.random.default_rng(0)
    n_img, per = 40, 5
    n_txt = n_img * per
    s_i2t, s_t2i = rng.standard_normal((n_img, n_txt)), rng.standard_normal((n_txt, n_img))
    img2txt = {i: list(range(i * per, (i + 1) * per)) for i in range(n_img)}
    txt2img = {j: j // per for j in range(n_txt)}
    for i in range(n_img):      # make the task non-trivial but solvable
        s_i2t[i, img2txt[i]] += 1.5
    for j in range(n_txt):
        s_t2i[j, txt2img[j]] += 1.5
    got = itm_eval(s_i2t, s_t2i, txt2img, img2txt)
    ranks = np.zeros(n_img)
    for index, score in enumerate(s_i2t):
        inds = np.argsort(score)[::-1]
        ranks[index] = min(np.where(inds == i)[0][0] for i in img2txt[index])
    tr = [100.0 * len(np.where(ranks < n)[0]) / len(ranks) for n in (1, 5, 10)]
    ranks = np.zeros(n_txt)
    for index, score in enumerate(s_t2i):
        inds = np.argsort(score)[::-1]
        ranks[index] = np.where(inds == txt2img[index])[0][0]
    ir = [100.0 * len(np.where(ranks < n)[0]) / len(ranks) for n in (1, 5, 10)]
    assert [got["txt_r1"], got["txt_r5"], got["txt_r10"]] == tr and [got["img_r1"], got["img_r5"], got["img_r10"]] == ir
    assert abs(got["r_mean"] - (sum(tr) / 3 + sum(ir) / 3) / 2) < 1e-12


def test_wgrad_split_k_fills_whole_waves():
    """blocks._split_k: the smallest factor that fills whole waves (<= 2 waves, >= 8 k-blocks per item)."""
    from xfm_b200.blocks import _split_k
    assert _split_k(18, 148, 296) == 8      # 768 x 768 outputs (128 x 256 tiles): 144 of 148 slots in one wave
    assert _split_k(72, 148, 296) == 2      # 768 x 3072
    assert _split_k(36, 74, 296) == 2       # 3072 x 768 on 74 SM pairs
    assert _split_k(27, 74, 296) == 5       # 2304 x 768: 135 of 148 slots over two waves beats 54 of 74 in one
    assert _split_k(591, 74, 23) == 1       # vocabulary-sized outputs never split
    assert _split_k(18, 148, 60) == 7       # short reductions: at least 8 k-blocks per item
    for tiles in (1, 3, 9, 18, 27, 36, 72, 100, 200):
        for kb in (4, 23, 60, 296):
            sk = _split_k(tiles, 148, kb)
            assert 1 <= sk <= max(1, kb // 8) and tiles * sk <= max(2 * 148, tiles)


def test_vqa_causal_bias_matches_the_reference_mask_rule():
    """xroberta.py:771-806 with is_decoder=True: position i sees keys j <= i that are not padding.  The product adds a
    causal [La, La] term and a key-padding term separately (-10000 each); after the softmax both forms give exact zeros."""
    from oracle import xfm_oracle as O
    atts = torch.tensor([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1]])
    ref = O.causal_extended_mask(atts)[:, 0]                                  # [B, L, L], one -10000 per masked pair
    L = atts.shape[1]
    ids = torch.arange(L)
    causal = (ids[None, :] > ids[:, None]).float() * -10000.0                  # model_generation._causal_bias
    key = (1.0 - atts.float()) * -10000.0                                      # RobertaStack.additive_mask
    mine = causal[None] + key[:, None, :]
    assert torch.equal(mine < 0, ref < 0)
    s = torch.randn(2, L, L)
    assert torch.equal(torch.softmax(s + mine, -1), torch.softmax(s + ref, -1))


def test_ctypes_struct_mirrors_match_the_header(tmp_path):
    """xfm_b200/lib.py mirrors xfm_gemm_params / xfm_attn_params field for field: sizes and the offset of every field are
    compared with what gcc computes from include/xfm_b200.h (a silent mismatch would shift every later pointer)."""
    import ctypes
    import os
    import shutil
    import subprocess

    from xfm_b200 import lib as L
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {"xfm_gemm_params": L.GemmParams, "xfm_attn_params": L.AttnParams}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "xfm_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for line in out:
        parts = line.split()
        if len(parts) != 3:
            continue
        cname, fname, val = parts
        cls = structs[cname]
        if fname == "size":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, fname).offset == int(val), (cname, fname, getattr(cls, fname).offset, val)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())
