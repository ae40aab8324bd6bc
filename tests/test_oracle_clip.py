"""CPU: oracle/clip_ref.py against HuggingFace's CLIPVisionModelWithProjection (architecture twin, random init).
openai/CLIP itself is not available offline: parity with it is unpinned (see oracle/clip_ref.py)."""
import pytest
import torch

from oracle import clip_ref


def _hf_from_params(p, cfg):
    tr = pytest.importorskip("transformers")
    hc = tr.CLIPVisionConfig(hidden_size=cfg["width"], intermediate_size=4 * cfg["width"], num_hidden_layers=cfg["layers"],
                             num_attention_heads=cfg["heads"], image_size=cfg["image_size"], patch_size=cfg["patch"],
                             projection_dim=cfg["out_dim"], hidden_act="quick_gelu", layer_norm_eps=1e-5)
    m = tr.CLIPVisionModelWithProjection(hc).eval()
    W = cfg["width"]
    sd = {"vision_model.embeddings.class_embedding": p["class_embedding"],
          "vision_model.embeddings.patch_embedding.weight": p["conv1.weight"],
          "vision_model.embeddings.position_embedding.weight": p["positional_embedding"],
          "vision_model.pre_layrnorm.weight": p["ln_pre.weight"], "vision_model.pre_layrnorm.bias": p["ln_pre.bias"],
          "vision_model.post_layernorm.weight": p["ln_post.weight"], "vision_model.post_layernorm.bias": p["ln_post.bias"],
          "visual_projection.weight": p["proj"].t().contiguous()}
    for i in range(cfg["layers"]):
        a, b = "transformer.resblocks.%d." % i, "vision_model.encoder.layers.%d." % i
        wq, wk, wv = p[a + "attn.in_proj_weight"].chunk(3, 0)
        bq, bk, bv = p[a + "attn.in_proj_bias"].chunk(3, 0)
        sd.update({b + "self_attn.q_proj.weight": wq, b + "self_attn.k_proj.weight": wk, b + "self_attn.v_proj.weight": wv,
                   b + "self_attn.q_proj.bias": bq, b + "self_attn.k_proj.bias": bk, b + "self_attn.v_proj.bias": bv,
                   b + "self_attn.out_proj.weight": p[a + "attn.out_proj.weight"], b + "self_attn.out_proj.bias": p[a + "attn.out_proj.bias"],
                   b + "layer_norm1.weight": p[a + "ln_1.weight"], b + "layer_norm1.bias": p[a + "ln_1.bias"],
                   b + "layer_norm2.weight": p[a + "ln_2.weight"], b + "layer_norm2.bias": p[a + "ln_2.bias"],
                   b + "mlp.fc1.weight": p[a + "mlp.c_fc.weight"], b + "mlp.fc1.bias": p[a + "mlp.c_fc.bias"],
                   b + "mlp.fc2.weight": p[a + "mlp.c_proj.weight"], b + "mlp.fc2.bias": p[a + "mlp.c_proj.bias"]})
    missing = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if "position_ids" not in k], missing
    return m


@pytest.mark.parametrize("name,B", [("tiny", 3), ("ViT-B/32", 1)])
def test_oracle_matches_hf_twin(name, B):
    cfg = clip_ref.CONFIGS[name]
    p = clip_ref.random_params(cfg, seed=1)
    m = _hf_from_params(p, cfg)
    torch.manual_seed(0)
    img = torch.randn(B, 3, cfg["image_size"], cfg["image_size"])
    with torch.no_grad():
        want = m(pixel_values=img).image_embeds
        got = clip_ref.encode_image(p, cfg, img)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), (got - want).abs().max()


def test_calc_matches_self_is_first():
    f = torch.nn.functional.normalize(torch.randn(40, 32), dim=-1)
    idx, val = clip_ref.calc_matches(f, 6)
    assert (idx[:, 0] == torch.arange(40)).all() and torch.allclose(val[:, 0], torch.ones(40), atol=1e-6)
