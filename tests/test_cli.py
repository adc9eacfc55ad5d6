"""bin/mcmc: the reference's command line (option table, messages, exit codes, stdout label line)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

CLI = os.path.join(ROOT, "bin", "mcmc")


@pytest.fixture(scope="module")
def cli(pkg):
    pkg.build.build_all()
    assert os.path.exists(CLI)
    return CLI


def run(cli, *args):
    p = subprocess.run([cli] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    return p.returncode, p.stdout, p.stderr


def write_edges(tmp_path, g, name="g.edgelist"):
    path = os.path.join(str(tmp_path), name)
    with open(path, "w") as f:
        for a, b in g["edges"]:
            f.write("%d\t%d\n" % (a, b))
    return path


def test_usage_and_argument_errors(cli):
    """exit 0 + usage on stderr with no arguments or -h; message + exit 1 on bad arguments
    (reference src/mcmc_main.cc:99-119, 154-218, 283-333)"""
    rc, out, err = run(cli)
    assert rc == 0 and out == "" and "Usage:" in err
    rc, out, err = run(cli, "-h")
    assert rc == 0 and "Usage:" in err
    rc, out, err = run(cli, "-y", 18, 14)
    assert rc == 1 and err == "edge_list_path is required (-e flag)\n"
    rc, out, err = run(cli, "-e", "x")
    assert rc == 1 and err == "types is required for bisbm mode (-y flag)\n"
    rc, out, err = run(cli, "-e", "x", "-y", 1, 2, 3)
    assert rc == 1 and "Number of types must be equal to 2!" in err
    rc, out, err = run(cli, "-e", "x", "-y", 18, 14, "-c", "exponential", "-a", 10, 1.5)
    assert rc == 1 and "alpha must be in ]0,1[" in err
    rc, out, err = run(cli, "-e", "x", "-y", 18, 14, "-c", "linear", "-a", 1, 2)
    assert rc == 1 and "eta must be in ]0, T_0]" in err
    rc, out, err = run(cli, "-e", "x", "-y", 18, 14, "-c", "bogus", "-a", 1)
    assert rc == 1 and "Invalid cooling schedule." in err
    rc, out, err = run(cli, "-e", "x", "-y", 18, 14)
    assert rc == 1 and "n is required (-n flag)" in err
    rc, out, err = run(cli, "-e", "x", "-y", 18, 14, "-n", 18, 14)
    assert rc == 1 and "number of partitions is required (-z flag)" in err
    rc, out, err = run(cli, "-e", "x", "-y", 18, 14, "-n", 10, 10, "-z", 1, 1)
    assert rc == 1 and "Types do not sum to the number of vertices!" in err
    rc, out, err = run(cli, "-e", "x", "-y", 18, 14, "-n", 18, 14, "-z", 1, 1, "--merge", "--chains", 4)
    assert rc == 1 and "agglomerative paths run one chain" in err
    rc, out, err = run(cli, "-e", "x", "-y", 18, 14, "-n", 18, 14, "-z", 1, 1, "--bogus")
    assert rc == 1 and "unrecognised option" in err


@pytest.mark.gpu
def test_travis_command_reproduces_reference_line(cli, tmp_path):
    """The CI / README maximisation command (reference .travis.yml:27) with the engine seeds pinned:
    stdout must be the reference's label line, stderr its three report lines."""
    g = load_golden("c1_seed1")
    path = write_edges(tmp_path, g)
    rc, out, err = run(cli, "-e", path, "-n", 4, 4, 4, 3, 3, 3, 3, 3, 3, 2, "-t", 1000, "-x", 100, "--maximize", "-c",
                       "exponential", "-a", 10, 0.1, "-y", 18, 14, "-z", 5, 5, "-E", 0.001, "--randomize", "--seed", 1,
                       "--gen_seed", 12345)
    assert rc == 0, err
    assert out == " ".join(str(x) for x in g["labels"]) + " \n"
    lines = err.strip().split("\n")
    assert lines[0] == "acceptance ratio %s" % float("%.6g" % g["accept"])
    assert lines[1] == "(Ka, Kb) = (5, 5) "
    assert lines[2] == "entropy: %s" % float("%.6g" % g["entropy"])


@pytest.mark.gpu
def test_readme_marginalisation_command_as_the_code_runs_it(cli, tmp_path):
    """README.md:87 -- the snapshot runs this as abrupt_cool annealing (-b/-f unused)."""
    g = load_golden("c2_abrupt")
    path = write_edges(tmp_path, g)
    rc, out, err = run(cli, "-e", path, "-n", *([50] * 20), "-t", 20000, "-x", 100, "-b", 1000, "-y", 500, 500, "-z", 10, 10,
                       "-f", 10, "-d", 1, "--gen_seed", 12345)
    assert rc == 0, err
    assert out == " ".join(str(x) for x in g["labels"]) + " \n"
    assert "acceptance ratio 0.0438125" in err and "entropy: 50026.1" in err


@pytest.mark.gpu
def test_membership_file_marginalize_and_estimate_modes(cli, tmp_path):
    g = load_golden("c2_const_k46")
    path = write_edges(tmp_path, g)
    mpath = os.path.join(str(tmp_path), "mb.txt")
    with open(mpath, "w") as f:
        f.write("\n".join(str(x) for x in g["labels0"]) + "\n")
    # membership file: -n / -z inferred, randomize forced off (reference src/mcmc_main.cc:247-278)
    rc, out, err = run(cli, "-e", path, "-y", 500, 500, "--membership_path", mpath, "-t", 3000, "-c", "constant", "-a", 1,
                       "--randomize", "-d", 3, "--gen_seed", 5)
    assert rc == 0, err
    assert "read membership from file" in err and "(Ka, Kb) = (4, 6)" in err
    lab = np.array(out.split(), dtype=np.int64)
    assert lab.size == 1000 and lab[:500].max() < 4 and lab[500:].min() >= 4 and lab.max() < 10
    # marginalisation over 32 chains: arg-max labels, one line
    rc, out, err = run(cli, "-e", path, "-y", 500, 500, "--membership_path", mpath, "--marginalize", "--chains", 32, "-b", 20,
                       "-t", 40, "-f", 4, "-d", 3)
    assert rc == 0, err
    lab = np.array(out.split(), dtype=np.int64)
    assert lab.size == 1000 and (lab[:500] < 4).all() and (lab[500:] >= 4).all()
    assert (lab == g["labels0"]).mean() > 0.8  # the planted partition is (close to) the marginal mode
    # estimate: CSV history sweep,Ka,Kb,loglik,labels...
    rc, out, err = run(cli, "-e", path, "-y", 500, 500, "--membership_path", mpath, "--estimate", "-b", 5, "-t", 30, "-f", 10)
    assert rc == 0, err
    rows = out.strip().split("\n")
    assert len(rows) == 3
    f0 = rows[0].split(",")
    assert f0[0] == "10" and f0[1] == "4" and f0[2] == "6" and float(f0[3]) < 0 and len(f0) == 4 + 1000
    # --uni: the 1D output format (sweep, K, loglik, labels)
    rc, out, err = run(cli, "-e", path, "-y", 500, 500, "--membership_path", mpath, "--estimate", "--uni", "-b", 5, "-t", 20, "-f", 10)
    assert rc == 0, err
    f0 = out.strip().split("\n")[0].split(",")
    assert f0[0] == "10" and f0[1] == "10" and float(f0[2]) < 0 and len(f0) == 3 + 1000
    # parallel restarts
    rc, out, err = run(cli, "-e", path, "-y", 500, 500, "-n", 125, 125, 125, 125, 83, 83, 83, 83, 84, 84, "-z", 4, 6, "--chains", 16,
                       "--randomize", "-t", 100000, "-x", 2000, "-c", "abrupt_cool", "-a", 50000, "-d", 1)
    assert rc == 0, err
    assert np.array(out.split()).size == 1000 and "entropy:" in err


@pytest.mark.gpu
def test_gpus_mode_shards_chains_and_all_reduces(cli, tmp_path):
    """--gpus N: chains dealt round-robin to N devices of the node by one process (graph replicated), the marginal
    histograms summed by ONE NCCL all-reduce behind the C ABI (bisbm_marginals_allreduce_local).  On a single-GPU box the
    second device does not exist and the command must fail cleanly; with two or more GPUs the label line must be a
    valid partition that agrees with the single-GPU marginals on most nodes."""
    import torch
    g = load_golden("c2_const_k46")
    path = write_edges(tmp_path, g)
    mb = os.path.join(str(tmp_path), "mb.txt")
    np.savetxt(mb, g["labels0"], fmt="%d")
    args = ["-e", path, "-y", 500, 500, "--membership_path", mb, "--marginalize", "-b", 30, "-t", 120, "-f", 4, "--chains", 64,
            "--seed", 5]
    rc1, out1, err1 = run(cli, *args)
    assert rc1 == 0, err1
    lab1 = np.array(out1.split(), dtype=np.int64)
    assert lab1.size == 1000
    rc2, out2, err2 = run(cli, *args, "--gpus", 2)
    if torch.cuda.device_count() < 2:
        assert rc2 == 1 and "device 1 out of range" in err2
        return
    assert rc2 == 0, err2
    lab2 = np.array(out2.split(), dtype=np.int64)
    assert lab2.size == 1000 and (lab2[:500] < 4).all() and (lab2[500:] >= 4).all() and (lab2[500:] < 10).all()
    assert (lab1 == lab2).mean() > 0.9
    # maximisation over restarts on two devices
    rc3, out3, err3 = run(cli, "-e", path, "-y", 500, 500, "--membership_path", mb, "-c", "abrupt_cool", "-a", 50000, "-t", 100000,
                          "-x", 10 ** 9, "--chains", 16, "--gpus", 2, "--seed", 3)
    assert rc3 == 0 and "entropy:" in err3 and len(out3.split()) == 1000


@pytest.mark.gpu
def test_merge_paths_reproduce_reference_lines(cli, tmp_path):
    """The agglomerative paths of main (reference src/mcmc_main.cc:350-451) with the engine seeds pinned: -g (singleton
    blocks merged down the sigma = 1.01 ladder to -z), -g -u (until a type has fewer than sqrt(2E)/2 blocks; prints Ka Kb
    first) and --mb labels with more blocks than -z.  stdout must be the reference's label line (tests/golden/merge.npz,
    generated from the unmodified reference build), stderr its two summary lines."""
    m = load_golden("merge")
    g = load_golden("c1_seed1")
    path = write_edges(tmp_path, g)
    common = ["-e", path, "-y", 18, 14, "-n", 18, 14, "-c", "abrupt_cool", "-a", 100, "-t", 320, "-x", 1000, "-d", 9,
              "--gen_seed", 12345]
    rc, out, err = run(cli, *common, "-z", 2, 3, "-g")
    assert rc == 0, err
    assert out == " ".join(str(x) for x in m["swg_labels"]) + " \n"
    assert "(Ka, Kb) = (2, 3) " in err and ("entropy: %s" % float("%.6g" % m["swg_entropy"])) in err
    rc, out, err = run(cli, *common, "-z", 2, 3, "-g", "-u")
    assert rc == 0, err
    ka, kb = (int(x) for x in m["swu_k"])
    assert out == "%d %d " % (ka, kb) + " ".join(str(x) for x in m["swu_labels"]) + " \n"
    assert "(Ka, Kb) = (%d, %d) " % (ka, kb) in err
    # labels with (8, 12) blocks, -z 4 6: the merge path for "more blocks than asked for"
    g2 = load_golden("c2_const_k46")
    path2 = write_edges(tmp_path, g2, "b.edgelist")
    KA, KB, p0, steps, await_ = m["b1000_path_args"]
    rc, out, err = run(cli, "-e", path2, "-y", 500, 500, "-n", 500, 500, "--mb", *[int(x) for x in m["b1000_labels0"]], "-z", int(KA),
                       int(KB), "-c", "abrupt_cool", "-a", int(p0), "-t", int(steps), "-x", int(await_), "-d", 11, "--gen_seed", 12345)
    assert rc == 0, err
    assert out == " ".join(str(x) for x in m["b1000_path_labels"]) + " \n"
    assert "(Ka, Kb) = (4, 6) " in err and ("entropy: %s" % float("%.6g" % m["b1000_path_entropy"])) in err
    # fewer blocks than asked for: the split path (property check only, see include/bisbm.h)
    rc, out, err = run(cli, "-e", path2, "-y", 500, 500, "-n", 500, 500, "--mb", *[int(x) for x in g2["labels0"]], "-z", 5, 6,
                       "-c", "abrupt_cool", "-a", 100, "-t", 2000, "-x", 1000, "-d", 2, "--gen_seed", 12345)
    assert rc == 0, err
    lab = np.array(out.split(), dtype=np.int64)
    assert "(Ka, Kb) = (5, 6) " in err and lab.size == 1000 and set(lab[:500]) == set(range(5)) and set(lab[500:]) == set(range(5, 11))
