"""MD loop throughput (development aid): the DHFR Langevin protocol and the ionic-fluid NVE run of bench.py, printed compactly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
for rep in range(2):
    d = bench.md_dhfr_block(torch, 0)
    print("md_dhfr: %.0f steps/s (%.1f us/step), updates %d, T %.1f K, E(1ps) %.1f" % (d["steps_per_s"], 1e3 * d["ms_per_step"], d["list_updates"], d["temperature_mean_K"], d["potential_energy_1ps"]))
i = bench.md_device_block(torch, 0)
for k in ("displacement_triggered", "update_every_10"):
    print("ionic %s: %.0f steps/s, updates %d, drift %.2e" % (k, i[k]["steps_per_s"], i[k]["list_updates"], i[k]["total_energy_drift_over_kinetic"]))
