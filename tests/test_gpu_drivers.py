"""Driver acceptance: the reference's own sapling_example.cpp, UNMODIFIED, compiled against our
drop-in include/sapling_api.h + libsapling_b200.so (oracle/_ref/sapling_example_b200, built by
`make -C oracle drivers_b200` where /root/reference exists) must report the same correctness line as the
reference binary built from the same source against the reference header (oracle/_ref/sapling_example)."""
import os
import re
import subprocess

import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

REFDIR = os.path.join(O.ROOT, "oracle", "_ref")


def _run(binary, fa, cwd, *args):
    p = subprocess.run([binary, fa, *args], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


@pytest.mark.skipif(not (os.path.exists(os.path.join(REFDIR, "sapling_example_b200")) and
                         os.path.exists(os.path.join(REFDIR, "sapling_example"))),
                    reason="reference drivers not prebuilt (oracle/_ref)")
def test_unmodified_sapling_example_runs_on_the_shim(tmp_path):
    g = O.synth_genome(O.SEED_G + 21, 300_000)
    outs = {}
    for tag, binary in (("ref", "sapling_example"), ("b200", "sapling_example_b200")):
        d = tmp_path / tag
        d.mkdir()
        fa = str(d / "g.fa")
        O.write_fasta(fa, g)
        outs[tag] = _run(os.path.join(REFDIR, binary), fa, str(d), "k=21", "nq=20000")
    pat = re.compile(r"Piecewise linear correctness: (\d+) out of (\d+)")
    a, b = pat.findall(outs["ref"]), pat.findall(outs["b200"])
    assert len(a) == 6 and a == b, (a, b)   # six query lengths k-10..k+80, same counts
    # both wrote the same index files
    for ext in (".sa", ".sap"):
        fa_ref, fa_b = str(tmp_path / "ref" / "g.fa") + ext, str(tmp_path / "b200" / "g.fa") + ext
        assert open(fa_ref, "rb").read() == open(fa_b, "rb").read(), ext
    # the queries both drivers generated (unseeded rand(), same sequence) are identical
    assert open(tmp_path / "ref" / "queries.out").read() == open(tmp_path / "b200" / "queries.out").read()
