cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "384 or flash or variants" > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest.log
for r in 1 2; do
VITCU_ATTN_KERNEL=solo timeout 300 python bench.py --img 384 --batch 64 --steps 10 --warmup 3 --no-extras >> gpurun_out/r02x_bench384_solo.json 2>>gpurun_out/r02x.err
timeout 300 python bench.py --img 384 --batch 64 --steps 10 --warmup 3 --no-extras >> gpurun_out/r02x_bench384_duo.json 2>>gpurun_out/r02x.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_flash_duo -s 13 -c 1 -o gpurun_out/r02x_flashduo python bench.py --img 384 --batch 64 --steps 2 --warmup 1 --no-extras > gpurun_out/r02x_ncu.log 2>&1
