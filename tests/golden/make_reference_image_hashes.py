#!/usr/bin/env python
"""tests/golden/reference_images.json: size and SHA-256 of the decoded RGB pixels of the PNGs the reference keeps under
doc/ for examples/density (README.md:201-206: the recorded output of the reference's Go binary on BASELINE configs[0]),
examples/tree-partition (README.md:65) and examples/nearest-neighbors (README.md:174-178).
The images themselves stay in the reference; the hashes travel (the GPU box has no /root/reference).
    python tests/golden/make_reference_image_hashes.py [/root/reference]
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["density_compare", "density_test", "density_test_periodic", "tree", "nearest_neighbours", "nearest_neighbours_periodic",
         "customColorMaps"]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = {}
    for name in NAMES:
        im = np.array(Image.open(os.path.join(ref, "doc", name + ".png")).convert("RGB"))
        out[name] = {"width": int(im.shape[1]), "height": int(im.shape[0]), "sha256_rgb": hashlib.sha256(im.tobytes()).hexdigest(),
                     "non_black_pixels": int((im.sum(2) > 0).sum()), "source": f"doc/{name}.png"}
    with open(os.path.join(HERE, "reference_images.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
