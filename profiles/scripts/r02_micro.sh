cd profiles/micro
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart shared -o /tmp/hash_insert hash_insert.cu 2>/dev/null
/tmp/hash_insert | tee /root/repo/gpurun_out/r02_hash_insert.txt
