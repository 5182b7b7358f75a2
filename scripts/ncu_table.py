"""Print key metrics of ncu reports as a markdown table.  usage: python scripts/ncu_table.py rep1.ncu-rep [rep2 ...]"""
import csv, subprocess, sys, io
W = [("gpu__time_duration.sum", "duration"), ("gpc__cycles_elapsed.avg.per_second", "SM clock"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
     ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("launch__registers_per_thread", "regs"), ("launch__block_size", "block"),
     ("launch__shared_mem_per_block_dynamic", "smem/CTA")]
print("| report | kernel | " + " | ".join(n for _, n in W) + " |")
print("|---|---|" + "---|" * len(W))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        cells = []
        for key, _ in W:
            v = ""
            for i, h in enumerate(hdr):
                if h == key:
                    try:
                        v = "%.4g %s" % (float(r[i]), units[i])
                    except ValueError:
                        v = r[i]
            cells.append(v.strip())
        name = r[hdr.index("Kernel Name")]
        name = name.replace("void ", "").split("(CUtensorMap")[0].split("(const")[0][:48]
        print("| %s | `%s` | %s |" % (rep.split("/")[-1].replace(".ncu-rep", ""), name, " | ".join(cells)))
