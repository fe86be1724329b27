import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g.load_package()
blobs = pkg.synth.model_blobs(os.path.join("/root/repo", "oracle", "_ref", "Network"))
imgs = pkg.synth.synthetic_images(64, 224, seed=3)
imgs = np.ascontiguousarray(imgs[np.arange(1024) % 64])
os.environ["VITB200_PRECISION"] = "bf16"
for persist in ("0", "1"):
    os.environ["VITB200_PERSIST"] = persist
    for rep in range(3):
        t = time.perf_counter(); pkg.vit_opencl(imgs, blobs); print(persist, rep, f"{1e3*(time.perf_counter()-t):.1f} ms", flush=True)
pkg.lib().vitb200_release_persistent()
