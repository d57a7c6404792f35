for o in "accel=1" "accel=2" "accel=2 --opt smem_nodes=0" "accel=2 --opt smem_nodes=600" "accel=1 --opt smem_nodes=1200" "accel=1 --opt smem_nodes=3900" "accel=1 --opt leaf_split=3" "accel=2 --opt leaf_split=3" "accel=1 --opt leaf_split=1"; do
echo "c4 $o: $(python tools/ab.py --spp 32 --size 3840 --opt config=c4 --opt pool_slots=16777216 --opt $o default)"
done
