set -e
cd $GRAFT_REPO_ROOT
PKG=pfe-raft-and-hyperprior-based-learned-video-compression_b200/lib
CUDART=$(python -c "import torch,os,glob; print(os.path.dirname(glob.glob(os.path.join(os.path.dirname(torch.__file__),'..','nvidia','cuda_runtime','lib','libcudart.so*'))[0]))" 2>/dev/null || echo /usr/local/cuda/lib64)
gcc -std=c99 -Iinclude examples/host_pair.c -o /tmp/host_pair -L$PKG -lrdvc_corr -lm -Wl,-rpath,$PWD/$PKG -Wl,--allow-shlib-undefined
LD_LIBRARY_PATH=/usr/local/cuda/lib64:$CUDART /tmp/host_pair
LD_LIBRARY_PATH=/usr/local/cuda/lib64:$CUDART /tmp/host_pair 136 240
