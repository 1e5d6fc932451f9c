#!/bin/bash
# SASS evidence for profiles/ (no GPU needed): which Blackwell mechanisms the built objects actually contain.
#   usage: bash tools/sass_extract.sh > profiles/r02_sass_extract.txt      (after `make -C ligero-prover_b200`)
B=ligero-prover_b200/build
count() { cuobjdump -sass $B/$1.o 2>/dev/null | grep -cE "$2"; }
echo "# SASS mnemonic counts per object (cuobjdump -sass, sm_100a), $(date -u +%F)"
echo "object,mnemonic,count,meaning"
echo "encode_8.o,UBLKCP.S.G,$(count encode_8 'UBLKCP\.S\.G'),TMA bulk copy global->shared of the message row (cp.async.bulk + mbarrier)"
echo "encode_8.o,SYNCS.ARRIVE.TRANS64,$(count encode_8 'SYNCS\.ARRIVE\.TRANS64'),mbarrier arrive.expect_tx for that copy"
echo "encode_8.o,SYNCS.PHASECHK.TRANS64.TRYWAIT,$(count encode_8 'SYNCS\.PHASECHK'),mbarrier try_wait"
echo "encode_8.o,IMAD.WIDE.U32(.X),$(count encode_8 'IMAD\.WIDE\.U32'),32x32->64 multiply-adds on carry chains (mad.lo.cc/madc.hi.cc pairs fuse)"
echo "encode_8.o,STG.E.ENL2.256,$(count encode_8 'STG\.E\.ENL2\.256'),one 256-bit store per codeword element (also the NVLink store when the sink is peer memory)"
echo "lat_ntt_kernel.o,IMAD.WIDE.U32(.X),$(count lat_ntt_kernel 'IMAD\.WIDE\.U32'),5 inlined Montgomery multiplications => ~146 wide multiply-adds each (136 products + address arithmetic)"
echo "ntt_10.o,IMAD.WIDE.U32(.X),$(count ntt_10 'IMAD\.WIDE\.U32'),fully unrolled 1024-point tile"
echo "ntt_10.o,LDG.E.ENL2.256,$(count ntt_10 'LDG\.E\.ENL2\.256'),256-bit element loads"
echo "sha_kernels.o,SHFL.BFLY,$(count sha_kernels 'SHFL\.BFLY'),lane-split hash chain: one butterfly exchange per round and direction"
echo "sha_kernels.o,SYNCS.EXCH.64,$(count sha_kernels 'SYNCS\.EXCH\.64'),mbarrier init of the K+W ring hand-over"
echo "sha_kernels.o,SYNCS.ARRIVE.TRANS64.A1T0,$(count sha_kernels 'SYNCS\.ARRIVE\.TRANS64\.A1T0'),ring slot hand-over between schedule warps and the chain warp"
echo "peer_kernels.o,STG.E.64.STRONG.SYS,$(count peer_kernels 'STG\.E\.64\.STRONG\.SYS'),st.release.sys of the round counter into peer memory"
echo "peer_kernels.o,LDG.E.64.STRONG.SYS,$(count peer_kernels 'LDG\.E\.64\.STRONG\.SYS'),ld.acquire.sys spin on the local flag array"
echo "peer_kernels.o,MEMBAR.*SYS,$(count peer_kernels 'MEMBAR\.[A-Z]+\.SYS'),system-scope fences around the hand-over"
echo "peer_kernels.o,NANOSLEEP,$(count peer_kernels 'NANOSLEEP'),back-off inside the bounded wait"
echo "cluster_encode_kernel.o,UCGABAR_ARV/WAIT,$(count cluster_encode_kernel 'UCGABAR'),cluster barriers around the distributed-shared-memory stages"
echo "cluster_encode_kernel.o,LD.E.128/ST.E.128,$(count cluster_encode_kernel '(LD|ST)\.E\.128'),generic 128-bit accesses to another CTA's shared memory (map_shared_rank)"
echo "eltwise_kernels.o,DFMA.RZ,$(count eltwise_kernels 'DFMA\.RZ'),FP64-pipe product halves of the alternative check_code sweep (LGR_COMBINE_CODE=dpf)"
echo "ubench_dpf.o,DFMA.RZ,$(count ubench_dpf 'DFMA\.RZ'),FP64-pipe Montgomery multiplication (measured alternative)"
echo "ubench_dpf.o,DADD,$(count ubench_dpf 'DADD'),its exponent-cancelling additions"
echo "all objects,HMMA|UTC.MMA|LDTM,$(cat $B/*.o | strings | grep -cE 'HMMA|UTC.MMA|LDTM'),no tensor-core instructions anywhere: the path is modular integer arithmetic"
