for l in 2 3; do for n in 900 1100 1300 1500; do
echo "accel=2 leaf_split=$l smem_nodes=$n: $(python tools/ab.py --spp 256 --opt accel=2 --opt leaf_split=$l --opt smem_nodes=$n default)"
done; done
