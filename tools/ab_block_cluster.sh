bash tools/quick_ab.sh r2_block
bash tools/quick_ab.sh r2_cluster -DNAFGPU_HUF_CLUSTER
