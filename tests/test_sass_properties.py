"""Static properties of the shipped library, checked on its SASS (no GPU needed): it is sm_100a code; the kernels whose
arithmetic must equal the reference's separate multiply-and-add hold no fused multiply-add outside
the IEEE divide / sqrt expansions; the volume records move with Blackwell's 256-bit loads and stores."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gunship_rs_b200", "libgunship_collision.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def kernels():
    if not os.path.exists(CUOBJDUMP) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is missing")
    txt = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    assert set(re.findall(r"arch = (sm_\w+)", txt)) == {"sm_100a"}
    out = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = f.split("\n", 1)
        out[name.strip()] = body
    return out


def _one(kernels, needle):
    hits = [b for n, b in kernels.items() if needle in n]
    assert len(hits) == 1, (needle, [n for n in kernels if needle in n])
    return hits[0]


def test_no_fused_multiply_add_where_the_reference_rounds_twice(kernels):
    # k_bvh_update (quaternion -> matrix, box AABB), k_pairs (integer filter) and k_narrow<false> (the three shape tests)
    # contain no division: any FFMA there would be a contracted a*b+c, i.e. a different rounding than the reference's
    for needle in ("12k_bvh_update", "7k_pairsILb0", "8k_narrowILb0"):
        assert not re.search(r"\bFFMA\b", _one(kernels, needle)), needle
    # kernels that divide (cell = floor(p / cell_size)) use FFMA only inside the IEEE-exact division sequence
    body = _one(kernels, "11k_cell_keys")
    assert re.search(r"MUFU\.RCP", body) and re.search(r"\bFCHK\b", body)


def test_volume_records_move_with_256_bit_accesses(kernels):
    assert len(re.findall(r"STG\.E[.A-Z0-9]*\.256", _one(kernels, "12k_bvh_update"))) >= 3    # rec + 2 per box record
    for needle in ("8k_narrowILb0", "11k_cell_keys", "20k_cell_table_permute"):
        assert re.search(r"LDG\.E[.A-Z0-9]*\.256", _one(kernels, needle)), needle


def test_radix_ranking_is_the_short_ballot_step(kernels):
    # 16 keys x (9 + 8 + 7) ballots for the three digit widths; the C++ form of the loop needed ~2x the LOP3 / ISETP
    body = _one(kernels, "17k_radix_downsweep")
    votes = len(re.findall(r"\bVOTE\.ANY\b", body))
    assert votes == 16 * (9 + 8 + 7)
    assert len(re.findall(r"\bISETP\b", body)) < votes // 2
