"""dmx_pack_reads against the reference's own pileup container (oracle/_ref/ref_pack: add_read into
map<snp, map<cell, map<umi, ...>>>, then the flattening walk) on the same read list: reads of a synth-v1 pileup
(C3 shape, SCALE of its barcodes), shuffled into a BAM-like arrival order, with 20 % duplicate UMI reads.
usage: python tools/pack_bench.py [SCALE=0.1]      (host only)"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demuxlet_b200 import api                      # noqa: E402
from demuxlet_b200.synth import make_config        # noqa: E402
from tests import packio                           # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
plp, _ = make_config("C3", scale=scale)
rng = np.random.default_rng(1)
E, R = plp.n_entries, plp.n_reads
ent_cell = np.repeat(np.arange(plp.ncells, dtype=np.int32), np.diff(plp.cell_ptr))
rd_ent = np.repeat(np.arange(E, dtype=np.int64), np.diff(plp.ent_ptr))
cell = ent_cell[rd_ent]
snp = plp.ent_snp[rd_ent]
allele = (plp.reads >> 8).astype(np.uint8)
bq = (plp.reads & 0xff).astype(np.uint8)
# 10-base UMIs, unique per read inside an entry; then 20 % extra reads repeating an existing (cell, snp, umi)
alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
umi_codes = alphabet[rng.integers(0, 4, size=(R, 10))]
dup = rng.integers(0, R, size=R // 5)
cell = np.concatenate([cell, cell[dup]]); snp = np.concatenate([snp, snp[dup]])
allele = np.concatenate([allele, rng.integers(0, 3, size=len(dup)).astype(np.uint8)])
bq = np.concatenate([bq, rng.integers(13, 41, size=len(dup)).astype(np.uint8)])
umi_codes = np.concatenate([umi_codes, umi_codes[dup]])
n = len(cell)
perm = np.argsort(snp + rng.random(n), kind="stable")      # BAM order: by position, cells interleaved
cell, snp, allele, bq, umi_codes = cell[perm], snp[perm], allele[perm], bq[perm], umi_codes[perm]
buf = np.ascontiguousarray(umi_codes).reshape(-1)
off = np.arange(n + 1, dtype=np.int64) * 10
t0 = time.perf_counter()
ours = api.pack_reads(cell, snp, (np.concatenate([buf, np.zeros(1, np.uint8)]), off), allele, bq, plp.ncells)
t_ours = time.perf_counter() - t0
out = {"n_reads": int(n), "ncells": int(plp.ncells), "entries": int(len(ours["ent_snp"])), "cores": os.cpu_count(),
       "ours_s": t_ours, "ours_reads_per_s": n / t_ours}
if os.path.exists(packio.REF_PACK):
    d = tempfile.mkdtemp()
    rds, csr = os.path.join(d, "a.rds"), os.path.join(d, "a.csr")
    with open(rds, "wb") as f:
        import struct
        f.write(b"DMXRDS01"); f.write(struct.pack("<3q", n, plp.ncells, plp.nsnps))
        for a in (cell.astype(np.int32), snp.astype(np.int32), allele, bq, off):
            f.write(np.ascontiguousarray(a).tobytes())
        f.write(buf.tobytes())
    r = packio.run_ref_pack(rds, csr)
    ref = packio.read_csr(csr)
    for k in ("cell_ptr", "ent_snp", "ent_ptr", "reads", "npass", "uniq"):
        assert np.array_equal(ours[k], ref[k]), k
    t_ref = r["t_add_read_s"] + r["t_flatten_s"]
    out.update({"reference_s": t_ref, "reference_reads_per_s": n / t_ref, "speedup": t_ref / t_ours, "identical": True})
print(json.dumps(out))
