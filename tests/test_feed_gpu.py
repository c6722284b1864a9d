"""Device half of the batch feeder (csrc/feed.cu through the C-ABI, xfm_b200.feed.DeviceFeeder) against the CPU transform it
replaces: transforms.ToTensor() + transforms.Normalize(mean, std) (+ hflip) of dataset/__init__.py:26-35, restated as
(u8 / 255 - mean) / std in fp32 (tests/test_feed_cpu.py pins that restatement to torchvision).  Bar: bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


from oracle.feed_oracle import pil_crop_resize, to_tensor_normalize as _cpu_transform  # noqa: E402  (the checkers)


@pytest.mark.parametrize("B,H,W", [(1, 16, 16), (5, 224, 224), (3, 384, 384), (2, 7, 12), (0, 224, 224)])
def test_image_u8_to_f32_is_bit_identical_to_the_cpu_transform(B, H, W):
    from xfm_b200 import feed, lib
    g = torch.Generator().manual_seed(B * 1000 + H)
    u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    if B:
        u8[0, 0, :4] = torch.tensor([[0, 127, 255], [255, 0, 1], [1, 2, 3], [254, 128, 64]], dtype=torch.uint8)
    n0 = lib.launch_count()
    out = lib.image_u8_to_f32(u8.cuda(), feed.CLIP_MEAN, feed.CLIP_STD)
    assert out.shape == (B, 3, H, W) and out.dtype == torch.float32
    assert torch.equal(out.cpu(), _cpu_transform(u8, feed.CLIP_MEAN, feed.CLIP_STD))
    assert lib.launch_count() - n0 == (1 if B else 0)
    if B:
        flip = (torch.arange(B) % 2 == 0).to(torch.uint8)
        out = lib.image_u8_to_f32(u8.cuda(), (0.5, 0.25, 0.125), (0.5, 2.0, 0.3), flip=flip.cuda())
        assert torch.equal(out.cpu(), _cpu_transform(u8, (0.5, 0.25, 0.125), (0.5, 2.0, 0.3), flip))


def test_image_u8_to_f32_every_byte_value_and_bad_arguments():
    from xfm_b200 import feed, lib
    u8 = torch.arange(256, dtype=torch.uint8).repeat(3)[: 4 * 64 * 3].reshape(1, 4, 64, 3).contiguous()
    out = lib.image_u8_to_f32(u8.cuda(), feed.CLIP_MEAN, feed.CLIP_STD)
    assert torch.equal(out.cpu(), _cpu_transform(u8, feed.CLIP_MEAN, feed.CLIP_STD))
    with pytest.raises(RuntimeError, match="W % 4"):
        lib.image_u8_to_f32(torch.zeros(1, 4, 6, 3, dtype=torch.uint8, device="cuda"), feed.CLIP_MEAN, feed.CLIP_STD)
    with pytest.raises(RuntimeError, match="std"):
        lib.image_u8_to_f32(torch.zeros(1, 4, 8, 3, dtype=torch.uint8, device="cuda"), feed.CLIP_MEAN, (1.0, 0.0, 1.0))


def test_device_feeder_delivers_the_reference_batches():
    """A loader that yields the reference's batch layout with uint8 crops -> the feeder's device batches equal the CPU-transformed
    batches moved with .cuda(), for list and dict batches, more batches than slots, ragged last batch, pinned and pageable
    inputs; H2D bytes are counted; the image is the fp32 NCHW tensor the model takes."""
    from xfm_b200 import feed
    g = torch.Generator().manual_seed(3)

    def host_batch(i, B, as_dict):
        u8 = torch.randint(0, 256, (B, 32, 48, 3), dtype=torch.uint8, generator=g)
        ids = torch.randint(0, 1000, (B, 12), generator=g)
        atts = (torch.rand(B, 12, generator=g) < 0.8).long()
        bbox = torch.rand(B, 4, generator=g)
        if i % 2:
            u8, ids = u8.pin_memory(), ids.pin_memory()
        return dict(image=u8, text_ids=ids, text_atts=atts, target_bbox=bbox, tag=i, none=None) if as_dict else \
            [u8, ids, atts, bbox, None]

    for as_dict in (False, True):
        batches = [host_batch(i, 6 if i < 6 else 3, as_dict) for i in range(7)]
        fd = feed.DeviceFeeder(batches, depth=2)
        seen = 0
        for host, dev in zip(batches, fd):
            hv = list(host.values()) if as_dict else host
            dv = list(dev.values()) if as_dict else dev
            if as_dict:
                assert list(dev.keys()) == list(host.keys()) and dev["tag"] == host["tag"] and dev["none"] is None
            assert dv[0].is_cuda and dv[0].dtype == torch.float32 and dv[0].shape == (hv[0].shape[0], 3, 32, 48)
            assert torch.equal(dv[0].cpu(), _cpu_transform(hv[0], feed.CLIP_MEAN, feed.CLIP_STD))
            for h, d in zip(hv[1:4], dv[1:4]):
                assert d.is_cuda and d.dtype == h.dtype and torch.equal(d.cpu(), h)
            seen += 1
        assert seen == 7
        want = sum(t.numel() * t.element_size() for b in batches for t in (b.values() if as_dict else b) if isinstance(t, torch.Tensor))
        assert fd.h2d_bytes == want

    # random hflip: each sample is either the transform or its mirror, both occur, and the draw is reproducible
    u8 = torch.randint(0, 256, (64, 8, 8, 3), dtype=torch.uint8, generator=g)
    outs = []
    for _ in range(2):
        (dev,) = list(feed.DeviceFeeder([[u8]], flip_prob=0.5, generator=torch.Generator().manual_seed(11)))
        outs.append(dev[0].cpu())
    assert torch.equal(outs[0], outs[1])
    plain = _cpu_transform(u8, feed.CLIP_MEAN, feed.CLIP_STD)
    same = (outs[0] == plain).flatten(1).all(1)
    mirrored = (outs[0] == plain.flip(3)).flatten(1).all(1)
    assert bool((same | mirrored).all()) and 8 < int(mirrored.sum()) < 56


def test_device_feeder_feeds_the_pretraining_model():
    """uint8 crops through the feeder -> XFM.forward: the losses equal those of the same batch transformed on the CPU."""
    from oracle import xfm_oracle as O
    from xfm_b200 import feed
    from xfm_b200.model_pretrain import XFM
    cfg = O.tiny_config(use_vision_tokenizer=False)
    B = 4
    model = XFM(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda:0").eval()
    batch = O.make_batch(cfg, B, L=24, M=6, seed=1, image_uniform=True)
    res = cfg["image_res"]
    u8 = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(2))
    keys = ("text_ids", "text_atts", "text_ids_masked", "masked_pos", "masked_ids")
    ineg, tneg = torch.roll(torch.arange(B), 1), torch.roll(torch.arange(B), -1)

    def run(image, b):
        model._forced_negatives = (ineg, tneg)
        out = model(image, b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                    masked_ids=b["masked_ids"], ret_mim_loss=False, data_source="image")
        return {k: float(v) for k, v in out.items() if k.startswith("loss_")}

    (dev,) = list(feed.DeviceFeeder([dict(image=u8, **{k: batch[k] for k in keys})]))
    with torch.no_grad():
        fed = run(dev["image"], dev)
        direct = run(_cpu_transform(u8, feed.CLIP_MEAN, feed.CLIP_STD).cuda(), {k: batch[k].cuda() for k in keys})
    assert fed.keys() == direct.keys() and fed["loss_itc"] > 0 and fed["loss_mlm"] > 0
    for k in fed:   # identical inputs; the loss reductions use fp32 atomics, hence not `==`
        assert abs(fed[k] - direct[k]) <= 1e-5 * max(1.0, abs(direct[k])), (k, fed[k], direct[k])


def test_device_tap_tables_equal_the_host_tables():
    """xfm_resize_taps (float64 on the GPU, explicitly rounded operations) against feed.pillow_bicubic_taps (numpy float64, pinned
    to PIL in tests/test_feed_cpu.py): every first tap, tap count and fixed-point tap identical, from 1-pixel crops to 16x
    down-scaling, both axes, ragged table widths."""
    import numpy as np
    from xfm_b200 import feed, lib
    rng = np.random.default_rng(5)
    for oh, ow in [(224, 224), (384, 384), (30, 50)]:
        sizes = [(int(h), int(w)) for h, w in zip(rng.integers(1, 3600, 70), rng.integers(1, 3600, 70))]
        sizes += [(oh, ow), (1, 1), (2 * oh, 2 * ow), (oh + 1, ow - 1), (3 * oh + 1, 5 * ow - 2)]
        plan = feed.crop_resize_plan(sizes, [(0, 0, w, h) for h, w in sizes], oh, ow)
        n0 = lib.launch_count()
        hb, hk, vb, vk = lib.resize_taps(plan["desc"].cuda(), oh, ow, plan["KH"], plan["KV"])
        assert lib.launch_count() - n0 == 1
        for name, got in zip(("hb", "hk", "vb", "vk"), (hb, hk, vb, vk)):
            assert torch.equal(got.cpu(), plan[name]), (name, oh, ow, int((got.cpu() != plan[name]).sum()))


@pytest.mark.parametrize("taps", ["device", "host"])
@pytest.mark.parametrize("res", [224, 384])
def test_crop_resize_is_bit_identical_to_pil(res, taps):
    """RandomResizedCrop / Resize with InterpolationMode.BICUBIC (dataset/__init__.py:28-30,63-67) on a ragged batch: the GPU
    result equals PIL's crop(box).resize((res, res), BICUBIC) byte for byte; chained with the normalize kernel it equals
    torchvision's Resize -> ToTensor -> Normalize."""
    import numpy as np
    from PIL import Image
    from xfm_b200 import feed, lib
    rng = np.random.default_rng(res)
    images, boxes = [], []
    for t in range(9):
        H, W = int(rng.integers(16, 640)), int(rng.integers(16, 640))
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        if t % 3 == 1:      # smooth content
            img = ((np.add.outer(np.arange(H) * 2, np.arange(W))[:, :, None] * np.array([1, 2, 3])) % 256).astype(np.uint8)
        x0, y0 = int(rng.integers(0, W - 8)), int(rng.integers(0, H - 8))
        box = (x0, y0, int(rng.integers(x0 + 4, W + 1)), int(rng.integers(y0 + 4, H + 1)))
        images.append(img)
        boxes.append(None if t % 4 == 0 else box)
    images.append(rng.integers(0, 256, size=(res, res, 3), dtype=np.uint8))     # already the target size: identity
    boxes.append(None)
    n0 = lib.launch_count()
    out = feed.crop_resize([torch.from_numpy(im) for im in images], boxes, res, res, taps=taps)
    assert lib.launch_count() - n0 == (3 if taps == "device" else 2)
    assert out.is_cuda and out.dtype == torch.uint8 and out.shape == (len(images), res, res, 3)
    got = out.cpu().numpy()
    for i, (img, box) in enumerate(zip(images, boxes)):
        ref = pil_crop_resize(img, box, res, res)
        assert (got[i] == ref).all(), (i, img.shape, box, int(np.abs(got[i].astype(int) - ref.astype(int)).max()))
    assert (got[-1] == images[-1]).all()
    tv = pytest.importorskip("torchvision.transforms")
    test_transform = tv.Compose([tv.Resize((res, res), interpolation=tv.InterpolationMode.BICUBIC), tv.ToTensor(),
                                 tv.Normalize(feed.CLIP_MEAN, feed.CLIP_STD)])
    whole = [i for i, b in enumerate(boxes) if b is None]
    f32 = lib.image_u8_to_f32(out, feed.CLIP_MEAN, feed.CLIP_STD).cpu()
    for i in whole:
        assert torch.equal(f32[i], test_transform(Image.fromarray(images[i])))
    assert feed.crop_resize([], [], res, res).shape == (0, res, res, 3)
