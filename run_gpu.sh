mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5
b() { python bench.py "$@" --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/last.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gs/s', round(d['roofline']['whole_step']['frac'],3), round(d['ms_per_step'],4), [(k['kind'], round(k['avg_ms'],4)) for k in d['roofline']['kernels']])" || tail -3 gpurun_out/last.err; }
echo -n "target: "; b --workload target
echo -n "target noWS: "; DSPB_NO_WS=1 b --workload target
for G in 8 16 32; do echo -n "config3 C=4096 G=$G: "; DSPB_FORCE_G=$G b --workload config3 --channels 4096; done
for G in 2 4 8; do echo -n "config3 C=1024 G=$G: "; DSPB_FORCE_G=$G b --workload config3; done
echo -n "config3 C=1024 G=4 noWS: "; DSPB_NO_WS=1 DSPB_FORCE_G=4 b --workload config3
echo -n "config5 C=1024: "; b --workload config5
