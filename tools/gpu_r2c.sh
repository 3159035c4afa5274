mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_i_mirrors.py -q -m gpu -x -p no:cacheprovider > gpurun_out/r2c_mirrors.log 2>&1; echo "mirrors exit $?"; tail -n 4 gpurun_out/r2c_mirrors.log
timeout 300 python tools/debug_gauss.py > gpurun_out/r2c_gauss.txt 2>&1; tail -n 14 gpurun_out/r2c_gauss.txt
timeout 1200 python -m pytest tests/test_gpu_e_segment.py tests/test_gpu_a_conv.py tests/test_gpu_b_unet.py -q -m gpu -p no:cacheprovider -s > gpurun_out/r2c_segment.log 2>&1; echo "segment exit $?"; grep -E "windows|passed|failed|Error" gpurun_out/r2c_segment.log | tail -n 20
timeout 900 python bench.py --workload cfg5 --steps 1 --warmup 1 --sweep-only 96:0.5,128:0.5,160:0.5,192:0.5,64:0.25 > gpurun_out/r2c_cfg5.json 2> gpurun_out/r2c_cfg5.err; echo "cfg5 exit $?"; tail -n 3 gpurun_out/r2c_cfg5.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c_cfg5.json"))
    for p in d["config"]["sweep"]:
        print(p["window"], p["overlap"], "Gvox/s", round(p["gvoxels_per_s"], 3), "convTF", round(p["conv_tflops_per_gpu"], 1), "frac", round(p["conv_frac_of_bf16_peak"], 3),
              "blend", p["blend_frac_of_hbm"], "fin", p["finalise_frac_of_hbm"], "ccl", p["ccl_frac_of_hbm"], p["stage_ms_max_rank"])
except Exception as e:
    print("no cfg5 json", e)
PY
