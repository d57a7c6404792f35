for a in 1 2; do for l in 2 3 4; do for b in 1024 768; do
echo "accel=$a leaf_split=$l trace_block=$b: $(python tools/ab.py --spp 256 --opt accel=$a --opt leaf_split=$l --opt trace_block=$b default)"
done; done; done
