#!/bin/bash
# tools/evidence_round.sh <tag>: everything profiles/ holds for a round, in one gpurun call (1 GPU): the ncu launch lists and --set full
# captures (tools/profile_round.sh), the bench lines of the default workload, the reference arm and the secondary workloads, the FMA-rate
# probe, and the compute-sanitizer runs.  Outputs land in gpurun_out/.
T=${1:-r02}
cd ${GRAFT_REPO_ROOT:-.}
python bench.py > gpurun_out/${T}_bench_default_1gpu.json 2> gpurun_out/${T}_bench_default_1gpu.err
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
for w in 1m_native_pan 256k_hq_room c4_1m_doppler_pan_room; do
  python bench.py --workload $w --no-cpu > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
bash tools/profile_round.sh $T > gpurun_out/${T}_profile_round.log 2>&1
( cd tools/probes && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_rate_probe fma_rate_probe.cu && ./fma_rate_probe ) > gpurun_out/${T}_fma_rate_probe.txt 2>&1
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_engine.py tests/test_polyphase.py -m gpu -q -k "not full_size and not c3_size and not 16k and not linearity" > gpurun_out/${T}_memcheck.log 2>&1
timeout 600 $SAN --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden and (c1_pan16 or hq16 or mixed_toggles or room_hq or room_pan_doppler or c3_hq_128tap)" > gpurun_out/${T}_racecheck.log 2>&1
tail -3 gpurun_out/${T}_memcheck.log gpurun_out/${T}_racecheck.log
ls -la gpurun_out | tail -40
