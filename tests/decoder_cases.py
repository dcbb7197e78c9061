"""Seeded decoder-head parity cases shared by the golden generator (reference, build container) and the tests."""

# name -> dict(kind, in_channels, inter, num_frames, n, h4, w4, + head options)
def case_table():
    c = {}
    c["emb_xyff_t8"] = dict(kind="embedding", in_channels=64, inter=[64, 64, 32, 32], num_frames=8, n=1, h4=24,
                            w4=32, embedding_size=4, dim_mode="xyff", tanh=True, seediness_output=True)
    c["emb_xyt_t8_batch2"] = dict(kind="embedding", in_channels=32, inter=[64, 32, 32, 64], num_frames=8, n=2,
                                  h4=24, w4=40, embedding_size=3, dim_mode="xyt", tanh=True, seediness_output=True)
    c["emb_xytff_t16"] = dict(kind="embedding", in_channels=32, inter=[32, 32, 32, 32], num_frames=16, n=1, h4=24,
                              w4=24, embedding_size=5, dim_mode="xytff", tanh=True, seediness_output=False)
    c["emb_ff_notanh_t4"] = dict(kind="embedding", in_channels=32, inter=[32, 64, 32, 32], num_frames=4, n=1,
                                 h4=32, w4=24, embedding_size=2, dim_mode="ff", tanh=False, seediness_output=True)
    c["emb_xyff_t2"] = dict(kind="embedding", in_channels=32, inter=[32, 32, 32, 32], num_frames=2, n=1, h4=24,
                            w4=48, embedding_size=4, dim_mode="xyff", tanh=True, seediness_output=False)
    c["seediness_t8"] = dict(kind="seediness", in_channels=64, inter=[64, 64, 32, 32], num_frames=8, n=1, h4=24,
                             w4=32)
    c["semseg_42_t8"] = dict(kind="semseg", in_channels=32, inter=[64, 64, 64, 64], num_frames=8, n=1, h4=24,
                             w4=32, num_out=42)
    c["semseg_4_t4"] = dict(kind="semseg", in_channels=32, inter=[32, 32, 32, 32], num_frames=4, n=1, h4=24,
                            w4=24, num_out=4)
    # the real channel widths of every shipped config (defaults.yaml:62), small spatial extent
    c["emb_fullwidth_t8"] = dict(kind="embedding", in_channels=256, inter=[256, 256, 128, 128], num_frames=8, n=1,
                                 h4=24, w4=32, embedding_size=4, dim_mode="xyff", tanh=True, seediness_output=False)
    c["seediness_fullwidth_t8"] = dict(kind="seediness", in_channels=256, inter=[256, 256, 128, 128], num_frames=8,
                                       n=1, h4=24, w4=32)
    # NUM_FRAMES 24 / 32 (three pooling slots like 16, common.py:22-23,34-35) and POOL_TYPE "max" (model_builder.py:30)
    c["emb_xyff_t24"] = dict(kind="embedding", in_channels=32, inter=[32, 32, 32, 32], num_frames=24, n=1, h4=24,
                             w4=24, embedding_size=4, dim_mode="xyff", tanh=True, seediness_output=True)
    c["seediness_t32"] = dict(kind="seediness", in_channels=32, inter=[32, 32, 32, 32], num_frames=32, n=1, h4=24,
                              w4=24)
    c["emb_xyff_maxpool_t8"] = dict(kind="embedding", in_channels=64, inter=[64, 64, 32, 32], num_frames=8, n=1,
                                    h4=24, w4=32, embedding_size=4, dim_mode="xyff", tanh=True, seediness_output=True,
                                    pool="max")
    c["semseg_4_maxpool_t16"] = dict(kind="semseg", in_channels=32, inter=[32, 32, 32, 32], num_frames=16, n=1, h4=24,
                                     w4=24, num_out=4, pool="max")
    return c


def case_seed(name):
    return sum(ord(ch) * (i + 1) for i, ch in enumerate(name)) % 100000


def build_case(name):
    """-> (state_dict, feature list in the head's expected order, case dict)."""
    from oracle import decoder_oracle as do
    case = case_table()[name]
    shapes = do.head_parameter_shapes(case["kind"], case["in_channels"], case["inter"],
                                      embedding_size=case.get("embedding_size"), dim_mode=case.get("dim_mode"),
                                      seediness_output=case.get("seediness_output", True),
                                      num_out=case.get("num_out"))
    sd = do.seeded_state_dict(shapes, case_seed(name))
    order = (4, 8, 16, 32) if case["kind"] == "semseg" else (32, 16, 8, 4)
    feats = do.seeded_features(case_seed(name) + 1, case["n"], case["in_channels"], case["num_frames"], case["h4"],
                               case["w4"], order=order)
    return sd, feats, case


def run_oracle(name, trace=None):
    from oracle import decoder_oracle as do
    sd, feats, case = build_case(name)
    pool = case.get("pool", "avg")
    if case["kind"] == "embedding":
        return do.embedding_head(sd, feats, case["num_frames"], case["embedding_size"], case["dim_mode"],
                                 case["tanh"], case["seediness_output"], trace=trace, pool_type=pool)
    if case["kind"] == "seediness":
        return do.seediness_head(sd, feats, case["num_frames"], trace=trace, pool_type=pool)
    return do.semseg_head(sd, feats, case["num_frames"], trace=trace, pool_type=pool)
