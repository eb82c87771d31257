#!/bin/bash
# GPU box: the evidence of one build in one call - full gpu test suite, bench line, ncu launch list of the bench command, DRAM traffic
# of one decode / encode step, ncu --set full of the encoder's two kernels, compute-sanitizer.  usage: scripts/final_capture.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 300 $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --verify 0 --enc-streams 1024 > $out/${tag}_bench_under_ncu.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:bp_ --csv --log-file $out/${tag}_traffic.csv python scripts/time_decode.py --streams 4096 --mode walk --distinct 1024 --reps 1 > $out/${tag}_traffic.log 2>&1
ncu --metrics $M --clock-control none -k "regex:ulcb_encode_kernel|ulcb_detect_kernel" --csv --log-file $out/${tag}_traffic_encode.csv python scripts/time_encode.py --streams 4096 --seconds 60 --reps 1 > $out/${tag}_traffic_encode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ulcb_encode_kernel -c 1 -o /tmp/${tag}_encode python scripts/time_encode.py --streams 296 --seconds 3 --reps 1 > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/${tag}_encode.ncu-rep --lines 40 > $out/${tag}_encode_ncu_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:ulcb_detect_kernel -c 1 -o /tmp/${tag}_detect python scripts/time_encode.py --streams 4096 --seconds 2 --reps 1 > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/${tag}_detect.ncu-rep --lines 30 > $out/${tag}_encode_detect_ncu_summary.txt 2>&1   # reports stay on the box: gpurun_out is capped at 64 MiB
{
echo "## memcheck"; compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py all 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_smoke|SUMMARY|Invalid|Error" | head -20
echo "## racecheck (encode: detector pre-pass + encode kernel)"; compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py encode 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_smoke|SUMMARY|hazard" | head -20
echo "## racecheck (encode, ULCB_ENC_PRE=0: detector warp inside the encode kernel)"; ULCB_ENC_PRE=0 compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py encode 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_smoke|SUMMARY|hazard" | head -20
echo "## synccheck (encode)"; compute-sanitizer --tool synccheck python scripts/sanitize_smoke.py encode 2>&1 | grep -v "Host Frame\|Device Frame\|=========$" | grep -E "COMPUTE-SANITIZER|sanitize_smoke|SUMMARY|error|Error|Barrier|at " | head -12
echo "## racecheck (decode)"; compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py decode 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_smoke|SUMMARY|hazard" | head -20
} > $out/${tag}_sanitizer.txt 2>&1
cat $out/${tag}_sanitizer.txt
