"""Turn an .ncu-rep of the decode/encode kernel into the text summaries kept under profiles/.
Usage: ncu_summarise.py rep tag source.cu nframes "description" [--traffic]"""
import csv, json, subprocess, sys, os
rep, tag, src, nframes, desc = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, u, v = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out, d = [desc], {}
for k in keep:
    if k in hdr:
        i = hdr.index(k); out.append("%s [%s] = %s" % (k, u[i], v[i])); d[k] = (float(v[i]), u[i])
open(os.path.join(root, "profiles", tag + "_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
srcsv = "/tmp/%s_src.csv" % tag
open(srcsv, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout)
br = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_lines.py"), srcsv, src, str(nframes)], capture_output=True, text=True).stdout
open(os.path.join(root, "profiles", tag + "_stage_breakdown.txt"), "w").write(br)
if "--traffic" in sys.argv:
    tr = sum(d[k][0] * mult[d[k][1]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    json.dump({"dram_bytes_per_frame": tr / nframes, "frames_in_capture": nframes, "dram_bytes_in_capture": tr,
               "algorithmic_bytes_per_frame": 14080, "source": "profiles/%s_ncu_summary.txt" % tag},
              open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(out[:12])); print(br.split("\n")[0])
