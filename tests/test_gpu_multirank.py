"""The programs under ``torchrun``: every rank owns one position range of the genome, tables are
completed by all-reduce, rank 0 writes — and the files equal the single-process ones, which equal the reference's
(tests/golden/ref_scripts).  Runs on ONE GPU: two / three ranks share cuda:0 and reduce through gloo
(``PB_DIST_BACKEND=gloo PB_DIST_ONE_DEVICE=1``); on a multi-GPU box the same command lines run over NCCL
(``tests/test_gpu_multirank.py::test_nccl_*``, skipped where fewer than two GPUs are visible)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import refgold as rg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bam(tmp_path_factory, cuda_device):
    from plastid_b200.bam_io import write_bam
    chrom_lengths, reads = rg.read_alignments()
    index = {c: i for i, c in enumerate(chrom_lengths)}
    path = str(tmp_path_factory.mktemp("multirank") / "reads.bam")
    write_bam(path, chrom_lengths, [(index[c], s, 16 if rev else 0, cig) for c, s, rev, cig in reads], record_aligned=True)
    return path


def torchrun(n, module, argv, one_device=True, port=29631):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    if one_device:
        env.update(PB_DIST_BACKEND="gloo", PB_DIST_ONE_DEVICE="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", module] + argv
    res = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    return res


ANN = ["--annotation_files", rg.inp("transcripts.bed"), "--annotation_format", "BED", "--bed_extra_columns", "gene_id"]
MSK = ["--mask_annotation_files", rg.inp("masks.bed"), "--mask_annotation_format", "BED"]


def cnt(bam):
    return ["--count_files", bam, "--countfile_format", "BAM", "--min_length", "25", "--max_length", "35"]


def body(path):
    with open(path) as fh:
        return [ln for ln in fh if not ln.startswith("##")]


def run_all(bam, tmp_path, n, one_device, sharding):
    sh = ["--sharding", sharding]
    out = str(tmp_path / "cir.txt")
    torchrun(n, "plastid_b200.bin.counts_in_region", [out] + cnt(bam) + ["--fiveprime", "--offset", "14"] + ANN + MSK + sh, one_device)
    assert body(out) == body(rg.out("counts_in_region_fiveprime14.txt"))
    base = str(tmp_path / "cs")
    torchrun(n, "plastid_b200.bin.cs", ["count", rg.out("cs_gene.positions"), base] + cnt(bam) + ["--threeprime", "--offset", "0"] + sh, one_device)
    assert body(base + ".txt") == body(rg.out("cs_count_threeprime.txt"))
    base = str(tmp_path / "mg")
    torchrun(n, "plastid_b200.bin.metagene", ["count", rg.out("mg_start_rois.txt"), base, "--min_counts", "5", "--normalize_over", "20", "80",
                                                "--fiveprime_variable", "--offset", rg.inp("p_offsets.txt")] + cnt(bam) + sh, one_device)
    gh, gr = rg.table(base + "_metagene_profile.txt")
    wh, wr = rg.table(rg.out("mg_start_median_metagene_profile.txt"))
    assert gh == wh
    rg.assert_float_columns_equal(gr, wr, {1}, rtol=1e-14, label="metagene")
    base = str(tmp_path / "mgm")
    torchrun(n, "plastid_b200.bin.metagene", ["count", rg.out("mg_start_rois.txt"), base, "--min_counts", "5", "--normalize_over", "20", "80",
                                                "--use_mean", "--fiveprime_variable", "--offset", rg.inp("p_offsets.txt")] + cnt(bam) + sh, one_device)
    gh, gr = rg.table(base + "_metagene_profile.txt")
    wh, wr = rg.table(rg.out("mg_start_mean_metagene_profile.txt"))
    assert gh == wh
    rg.assert_float_columns_equal(gr, wr, {1}, rtol=1e-12, label="metagene mean")      # column sums are added rank by rank
    base = str(tmp_path / "mgc")
    torchrun(n, "plastid_b200.bin.metagene", ["count", rg.out("mg_stop_rois.txt"), base, "--min_counts", "5", "--normalize_over", "-80", "-20",
                                                "--center", "--nibble", "10"] + cnt(bam) + sh, one_device)
    gh, gr = rg.table(base + "_metagene_profile.txt")
    wh, wr = rg.table(rg.out("mg_stop_center_metagene_profile.txt"))
    rg.assert_float_columns_equal(gr, wr, {1}, rtol=1e-9, label="metagene center")
    base = str(tmp_path / "ps")
    torchrun(n, "plastid_b200.bin.psite", [rg.out("mg_start_rois.txt"), base, "--min_counts", "5", "--normalize_over", "20", "80",
                                             "--require_upstream"] + cnt(bam) + sh, one_device)
    assert body(base + "_p_offsets.txt") == body(rg.out("psite_median_p_offsets.txt"))
    gh, gr = rg.table(base + "_metagene_profiles.txt")
    wh, wr = rg.table(rg.out("psite_median_metagene_profiles.txt"))
    rg.assert_float_columns_equal(gr, wr, set(range(1, len(wh))), rtol=1e-14, label="psite")
    base = str(tmp_path / "ph")
    torchrun(n, "plastid_b200.bin.phase_by_size", [rg.out("mg_start_rois.txt"), base, "--codon_buffer", "5"] + cnt(bam)
             + ["--fiveprime", "--offset", "14"] + sh, one_device)
    assert body(base + "_phasing.txt") == body(rg.out("phase_phasing.txt"))
    base = str(tmp_path / "wig")
    torchrun(n, "plastid_b200.bin.make_wiggle", ["-o", base, "--output_format", "variable_step", "--fiveprime", "--offset", "14"] + cnt(bam) + sh,
             one_device)
    for suffix in ("fw", "rc"):
        with open("%s_%s.wig" % (base, suffix)) as fh:
            got = [ln.rstrip("\n") for ln in fh if not ln.startswith("track")]
        assert got == rg.track_lines(rg.out("wig_fiveprime14_%s.wig.gz" % suffix))
    folder = str(tmp_path / "cv")
    torchrun(n, "plastid_b200.bin.get_count_vectors", [folder] + cnt(bam) + ["--fiveprime", "--offset", "14"] + ANN + MSK
             + ["--out_prefix", "cv_", "--format", "%d"] + sh, one_device)
    want = rg.count_vectors()
    for name in list(want)[::9]:
        assert np.array_equal(np.loadtxt(os.path.join(folder, name), ndmin=1), want[name]), name


@pytest.mark.parametrize("n,sharding", [(2, "positions"), (3, "chromosomes")])
def test_programs_under_torchrun_equal_the_reference_outputs(bam, tmp_path, n, sharding):
    run_all(bam, tmp_path, n, True, sharding)


def test_nccl_programs_on_two_gpus(bam, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`)")
    run_all(bam, tmp_path, 2, False, "positions")
