#!/usr/bin/env python
"""Does running two half-batches on two streams beat one full batch?  (LayerNorm CTAs of one lane can share
the SMs with the persistent GEMM CTAs of the other.)  Two engines, two host threads, same GPU."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
blobs = pkg.synth.model_blobs(None, 224, seed=7)
steps = 20


def run(batch, lanes):
    engs = []
    for i in range(lanes):
        e = pkg.Engine(0, 224, pkg.BF16, max_batch=batch)
        e.load_weights(blobs)
        e.stage(pkg.synth.synthetic_images(batch, 224, seed=i))
        for _ in range(3):
            e.forward_resident(batch)
        engs.append(e)
    res = [0.0] * lanes

    def work(i):
        res[i] = engs[i].time_resident(batch, steps)

    for rep in range(3):
        ths = [threading.Thread(target=work, args=(i,)) for i in range(lanes)]
        pkg.lib().vitcu_device_sync()
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        wall = time.perf_counter() - t0
        print(f"batch {batch} x {lanes} lane(s): {lanes * batch * steps / wall:.0f} images/s (wall), per-lane device ms/step {[round(r / steps, 3) for r in res]}")
    for e in engs:
        e.close()


run(256, 1)
run(128, 2)
run(256, 2)
run(128, 1)
