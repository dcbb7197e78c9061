"""Generate tests/golden/writeback_golden.npz with the UNMODIFIED reference DavisOutputGenerator (CPU).

Run in the build container only:  python tests/golden/gen_writeback_golden.py
Feeds the track labels of tests/golden/chain_golden.npz (the reference OnlineChainer's output) through
DavisOutputGenerator.process_sequence (davis.py:38-112), reads the PNGs it writes back, checks that
oracle/writeback_oracle.py reproduces them exactly and that no interpolated value is within 1e-6 of the 0.5
threshold (except exact ties), and stores the id maps.
"""
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402

_refshim.install()
import torch  # noqa: E402
from PIL import Image  # noqa: E402
from stemseg.config import cfg  # noqa: E402
from stemseg.inference.output_utils.davis import DavisOutputGenerator  # noqa: E402

from chain_cases import CASES, make_video  # noqa: E402
from oracle import gather_oracle as go  # noqa: E402
from oracle import writeback_oracle as wo  # noqa: E402
from writeback_cases import WRITEBACK_CASES  # noqa: E402


def main():
    chain = np.load(os.path.join(HERE, "chain_golden.npz"))
    out = {}
    for name, wb in WRITEBACK_CASES.items():
        masks, _ = make_video(**CASES[wb["video"]])
        t_total, h, w = masks.shape
        coords, _ = go.masks_to_coord_list(masks.astype(bool))
        labels = [chain["%s/track/%d" % (wb["video"], t)].astype(np.int64) for t in range(t_total)]
        ids = chain[wb["video"] + "/ids"].tolist()
        lifetimes = dict(zip(ids, chain[wb["video"] + "/lifetimes"].tolist()))
        # dict order of the reference = first appearance (frame-major, ids ascending per frame): rebuild it
        order = []
        for lab in labels:
            for i in np.unique(lab).tolist():
                if i not in order:
                    order.append(i)
        lifetimes = {i: lifetimes[i] for i in order}
        pt_counts = dict(zip(ids, chain[wb["video"] + "/pt_counts"].tolist()))
        cfg.INPUT.update_param("MIN_DIM", wb["min_dim"])
        cfg.INPUT.update_param("MAX_DIM", wb["max_dim"])
        image_h, image_w = wb["image_dims"]
        with tempfile.TemporaryDirectory() as tmp:
            gen = DavisOutputGenerator(tmp, -1, False, upscaled_inputs=False)
            seq = SimpleNamespace(image_dims=(image_h, image_w), id="seq")
            keep, _ = gen.process_sequence(
                seq, [(torch.from_numpy(y), torch.from_numpy(x)) for y, x in coords],
                [torch.from_numpy(l) for l in labels], pt_counts, lifetimes, None, (h, w), 4.0, wb["max_tracks"],
                device="cpu")
            ref = np.stack([np.array(Image.open(os.path.join(tmp, "results", "seq", "%05d.png" % t)))
                            for t in range(t_total)], 0).astype(np.uint8)
        o_keep = wo.instances_to_keep(lifetimes, -1, wb["max_tracks"])
        assert o_keep == keep, (o_keep, keep)
        ora = wo.id_maps(coords, labels, o_keep, (h, w), 4.0, (image_h, image_w), wb["min_dim"], wb["max_dim"])
        assert np.array_equal(ref, ora), "%s: oracle differs at %d pixels" % (name, int((ref != ora).sum()))
        margin = wo.threshold_margin(coords, labels, o_keep, (h, w), 4.0, (image_h, image_w), wb["min_dim"],
                                     wb["max_dim"])
        assert margin > 1e-6, "%s: a value sits within %g of the threshold" % (name, margin)
        out[name + "/maps"] = ref
        out[name + "/keep"] = np.array(keep, np.int64)
        print(name, ref.shape, "instances", keep, "margin %.2e" % margin, "nonzero", int((ref > 0).sum()))
    path = os.path.join(HERE, "writeback_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
