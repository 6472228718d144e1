"""TEST INFRASTRUCTURE ONLY.  Parameter names and shapes of ``model_zigma.ZigMa`` for a constructor config, restated from the
reference's constructors so that the CPU baseline can synthesise a state dict without instantiating any model class:

  ZigMa.__init__                         /root/reference/model_zigma.py:560-700 (embedders, pos_embed, blocks, norm_f, final layer)
  Mamba.__init__ (v1 / v2 / zigzag ...)  /root/reference/dis_mamba/mamba_ssm/modules/mamba_simple.py:63-226
  Block / FinalLayer / TimestepEmbedder  /root/reference/model_zigma.py:404-414, 322-340, 228-246

tests/test_oracle_shapes.py checks the result against the product model's ``state_dict()`` for every benchmark config and
against the shape tables of the golden fixtures (which come from the UNMODIFIED reference model)."""
import math


def zigma_state_shapes(cfg):
    D, depth = cfg["embed_dim"], cfg["depth"]
    C, p = cfg["in_channels"], cfg.get("patch_size", 1)
    vf = cfg.get("video_frames", 0)
    num_patches = (cfg["img_dim"] // p) ** 2
    sc = cfg.get("ssm_cfg") or {}
    N, W, expand = sc.get("d_state", 16), sc.get("d_conv", 4), sc.get("expand", 2)
    E = expand * D
    R = math.ceil(D / 16) if sc.get("dt_rank", "auto") == "auto" else sc["dt_rank"]
    s = {"x_embedder.proj.weight": (D, C, p, p), "x_embedder.proj.bias": (D,),
         "t_embedder.mlp.0.weight": (D, 256), "t_embedder.mlp.0.bias": (D,),
         "t_embedder.mlp.2.weight": (D, D), "t_embedder.mlp.2.bias": (D,)}
    if cfg.get("use_pe", 0) in (1, 2):
        s["pos_embed"] = (1, num_patches * max(vf, 1), D)
    if cfg.get("tpe", False):
        s["temporal_pos_embedding"] = (1, vf, D)
    if cfg.get("has_text", False):
        s["y_embedder.weight"], s["y_embedder.bias"] = (D, cfg["d_context"]), (D,)
    elif cfg.get("num_classes", -1) > 0:
        s["y_embedder.embedding_table.weight"] = (cfg["num_classes"], D)      # dropout_prob = 0: no extra "null" row
    two = cfg.get("scan_type", "v2") == "v2"
    for i in range(depth):
        m = f"blocks.{i}.mixer."
        s[m + "A_log"], s[m + "D"] = (E, N), (E,)
        if two:
            s[m + "A_b_log"], s[m + "D_b"] = (E, N), (E,)
        s[m + "in_proj.weight"] = (2 * E, D)
        s[m + "conv1d.weight"], s[m + "conv1d.bias"] = (E, 1, W), (E,)
        s[m + "x_proj.weight"] = (R + 2 * N, E)
        s[m + "dt_proj.weight"], s[m + "dt_proj.bias"] = (E, R), (E,)
        if two:
            s[m + "conv1d_b.weight"], s[m + "conv1d_b.bias"] = (E, 1, W), (E,)
            s[m + "x_proj_b.weight"] = (R + 2 * N, E)
            s[m + "dt_proj_b.weight"], s[m + "dt_proj_b.bias"] = (E, R), (E,)
        s[m + "out_proj.weight"] = (D, E)
        s[f"blocks.{i}.norm.weight"] = (D,)
        if not cfg.get("rms_norm", True):
            s[f"blocks.{i}.norm.bias"] = (D,)
        if cfg.get("has_text", False):      # CrossAttention(query_dim=D, context_dim=D, heads=8, dim_head=64): model_zigma.py:382-386
            a = f"blocks.{i}.msa."
            s[a + "to_q.weight"], s[a + "to_k.weight"], s[a + "to_v.weight"] = (512, D), (512, D), (512, D)
            s[a + "to_out.0.weight"], s[a + "to_out.0.bias"] = (D, 512), (D,)
        nmod = 6 if cfg.get("has_text", False) else 3
        s[f"blocks.{i}.adaLN_modulation.1.weight"], s[f"blocks.{i}.adaLN_modulation.1.bias"] = (nmod * D, D), (nmod * D,)
    s["norm_f.weight"] = (D,)
    if not cfg.get("rms_norm", True):
        s["norm_f.bias"] = (D,)
    s["final_layer.linear.weight"], s["final_layer.linear.bias"] = (p * p * C, D), (p * p * C,)
    return s
