# instrumented warp-specialised builds: per-warp clock split of block 0 (one decode of nyx512)
for tag in "$@"; do
cp vizaly-foresight_b200/libzfb_$tag.so vizaly-foresight_b200/libzfb.so
echo "== $tag"
timeout 300 python bench.py --workload nyx512 --steps 1 --warmup 0 --no-e2e --no-cpu 2>&1 | grep -E "^(P|T)[0-9]+ " | tail -24
done
