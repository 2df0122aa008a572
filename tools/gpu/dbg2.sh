mkdir -p gpurun_out
for lib in - tools/gpu/var/libantiz_A.so tools/gpu/var/libantiz_B.so; do for w in c1 c2s; do python tools/gpu/dbg2.py $lib $w; done; done 2>&1 | grep -v Warning | tee gpurun_out/dbg2.log
