"""Text summary of one `ncu --set full --import-source on` capture, as committed under profiles/.
usage: report.py report.ncu-rep natoms "command line that was profiled" ["comment"]"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
rep, natoms, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
print(f"# {cmd}")
if len(sys.argv) > 4:
    print(f"# {sys.argv[4]}")
print("# (numbers under the profiler are for the kernel's SHARE and pipe mix, not bench values)\n")
print("## instruction mix / stalls (aggregated from --page source and --page raw; per TR, all warps of one atom)")
print(subprocess.run([sys.executable, os.path.join(here, "summarize_ncu.py"), rep, natoms], capture_output=True, text=True).stdout)
print("## hottest source lines (share of executed warp instructions / of stall samples)")
print(subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, "30"], capture_output=True, text=True).stdout)
print("## ncu --page details")
print(subprocess.run(f"ncu -i {rep} --page details", shell=True, capture_output=True, text=True).stdout)
