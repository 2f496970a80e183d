echo shift2; python bench.py --kernel-only --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-300
echo shift16; X360_LIB=$PWD/360extractor_b200/lib/libx360_m5.so python bench.py --kernel-only --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-300
