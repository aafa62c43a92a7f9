import re,csv,collections,sys
sass, src, raw = sys.argv[1:4]
exec(open('tools/sass_lines.py').read().split("tot=sum")[0].replace("sass, ncucsv = sys.argv[1], sys.argv[2]","sass, ncucsv = sys.argv[1], sys.argv[2]").replace("topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40","topn=40"))
lines=open('your-voice-tts_b200/csrc/frame_kernels.cuh').read().split('\n')
def find(s): 
    return next(i+1 for i,l in enumerate(lines) if s in l)
marks=[('setup',0),('stage helpers',find('span staging helpers')),('tile loop/stage',find('previous segment is done with planes')),('frame load',find('window the frame')),
       ('middle',find('per-bin step -> conj')),('shfl2',find('hand the partner its half')),('transform glue',find('1024-point transform, 32 x 32')),('analysis out',find('spectrum out')),
       ('slot store',find("the warp's overlap-add slot")),('post-frame sync/stage_load',find('next span: its loads are issued by each warp')),('OLA',find('overlap-add + window + 1/(N wss) + store')),('stage_store/sync',find('if (have_next && !(a.debug & 2)) stage_store(')),('sc',find('if constexpr (MODE == MODE_GL_ITER && SC)'))]
helper_end=find('// the kernel')
def region(l):
    if l is None: return 'none'
    f,n=l
    if f in ('fft32.cuh','fft32p.cuh'): return 'fft32'
    if f!='frame_kernels.cuh': return f
    if n<helper_end: return 'passes(twiddle/exchange/cp.async)'
    r='setup'
    for name,start in marks:
        if n>=start: r=name
    return r
reg=collections.Counter(); regs=collections.Counter()
for l,c in agg.items(): reg[region(l)]+=c; regs[region(l)]+=sm[l]
tot=sum(reg.values()); ts=sum(regs.values())
nfr=32000
print('total warp-instr',tot,'per frame',tot/nfr)
for r,c in reg.most_common(): print(f"{r:36s} {c:12d} {100*c/tot:5.1f}% samples {100*regs[r]/ts:5.1f}%  per-frame {c/nfr:.0f}")
rows=list(csv.reader(open(raw))); hdr,units,data=rows[0],rows[1],rows[2:]
for w in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.max','lts__t_bytes.sum','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']:
    if w in hdr: print(w, units[hdr.index(w)], [d[hdr.index(w)] for d in data])
