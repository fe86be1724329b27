set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bench_config_parity.py -m gpu -x -q -s -k "fp8" > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
python - > gpurun_out/r02n_fp8_bench.log 2>&1 <<'PY'
import sys, os, json, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
pkg = g.load_package()
blobs = pkg.synth.model_blobs(os.path.join('oracle','_ref','Network'), 224, seed=0)
imgs = pkg.synth.synthetic_images(256, 224, seed=1234)
for prec, name in ((pkg.BF16, 'bf16'), (pkg.FP8, 'fp8'), (pkg.BF16, 'bf16'), (pkg.FP8, 'fp8')):
    with pkg.Engine(0, 224, prec, max_batch=256) as eng:
        eng.load_weights(blobs)
        eng.stage(imgs)
        for _ in range(5): eng.forward_resident(256)
        ms = eng.time_resident(256, 20) / 20
        tl = eng.profile_timeline(256, 2)
        gm, gl = eng.profile_gemms(256, 3)
        print(json.dumps({"precision": name, "ms_per_step": ms, "images_per_s": 256e3/ms, "timeline": tl, "gemm_ms": gm, "gemm_launches": gl}))
PY
