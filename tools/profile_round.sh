#!/bin/bash
# tools/profile_round.sh <tag>: the ncu evidence of a round, run on the GPU box (gpurun): launch list of the default workload
# + one `--set full` capture per hot kernel at its workload's full size.  Reports land in gpurun_out/; summaries are written
# afterwards with tools/ncu_summary.py into profiles/.
T=${1:-r02}
NCU="ncu --clock-control none"
B="python bench.py --no-cpu --no-also --pipeline 1 --blocks 1 --steps 6 --warmup 3"
$NCU --metrics gpu__time_duration.sum -k regex:'k1_|k2_|k3_|k4|k_p|k_r|k_a' -c 60 --csv --log-file gpurun_out/${T}_launches_c4.csv $B > gpurun_out/${T}_launches_c4.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:'k1_|k2_|k3_|k4|k_p|k_r|k_a' -c 60 --csv --log-file gpurun_out/${T}_launches_c3.csv $B --workload c3_65536_hq_128tap > gpurun_out/${T}_launches_c3.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:'k1_|k2_|k3_|k4|k_p|k_r|k_a' -c 80 --csv --log-file gpurun_out/${T}_launches_room.csv $B --workload 256k_hq_room > gpurun_out/${T}_launches_room.log 2>&1
F="$NCU --set full --import-source on -c 1"
$F -k regex:k2_mix_stereo -s 4 -o gpurun_out/${T}_k2_stereo_c4 -f $B > /dev/null 2>&1
$F -k regex:k1_params -s 4 -o gpurun_out/${T}_k1_c4 -f $B > /dev/null 2>&1
$F -k regex:k2_mix_ambisonic -s 4 -o gpurun_out/${T}_k2_sh_c3 -f $B --workload c3_65536_hq_128tap > /dev/null 2>&1
$F -k regex:k2_mix_ambisonic -s 4 -o gpurun_out/${T}_k2_sh_256k -f $B --workload 256k_hq > /dev/null 2>&1
$F -k regex:k2_mix_ambisonic -s 4 -o gpurun_out/${T}_k2_sh_c5 -f $B --workload c5_voice_ambient > /dev/null 2>&1
$F -k regex:k_room -s 4 -o gpurun_out/${T}_k_room -f $B --workload 256k_hq_room > /dev/null 2>&1
$F -k regex:k4a_decode -s 4 -o gpurun_out/${T}_k4a_c3 -f $B --workload c3_65536_hq_128tap > /dev/null 2>&1
ls -la gpurun_out/${T}_*
