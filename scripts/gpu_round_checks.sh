#!/bin/bash
# GPU tests, the real-image report and the bench lines of the other BASELINE configurations (1 GPU).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/c_pytest.log
tail -4 gpurun_out/c_pytest.log
timeout 300 python scripts/real_image_report.py > gpurun_out/real_images_r1.md 2> gpurun_out/real_images_r1.err
cat gpurun_out/real_images_r1.md; tail -n 3 gpurun_out/real_images_r1.err
# configs[4]: 64 frames x 4000 features against the 1 M database
timeout 600 python bench.py --features 4000 --steps 5 --no-cpu-baseline > gpurun_out/c_bench_cfg4.json 2> gpurun_out/c_bench_cfg4.err
# configs[1]: 100-object DB, single 2000-feature frame
timeout 600 python bench.py --objects 100 --frames 1 --lanes 1 --pose-warps 8 --steps 50 > gpurun_out/c_bench_cfg1.json 2> gpurun_out/c_bench_cfg1.err
cat gpurun_out/c_bench_cfg4.json gpurun_out/c_bench_cfg1.json
tail -n 2 gpurun_out/c_bench_cfg4.err gpurun_out/c_bench_cfg1.err
