mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_chain2 -s 9 -c 1 -f -o gpurun_out/r2o_chain_src python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/r2o_ncu_a.log 2>&1; echo "chain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:token_fused -s 1 -c 1 -f -o gpurun_out/r2o_token_src python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/r2o_ncu_b.log 2>&1; echo "token rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sig_attention -s 8 -c 1 -f -o gpurun_out/r2o_attn_src python bench.py --profile-only --steps 1 --warmup 1 > gpurun_out/r2o_ncu_c.log 2>&1; echo "attn rc=$?"
ls -la gpurun_out; du -sm gpurun_out
