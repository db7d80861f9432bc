"""The delimited-text readers (src/readData.jl:85-96, 159-165) through the C-ABI, host mode (needs no GPU): the
library's parser against the oracle's line-by-line restatement on files written in the BXD layouts the
reference's README describes (header row, id column, sex column / complement probability columns)."""
import os

import numpy as np
import pytest

import blmm_oracle as orc
from blmm_b200 import BlmmError, readBXDgeno, readBXDpheno, readGenoProb_ExcludeComplements, read_csv_matrix


def write_pheno(path, n=79, m=53, seed=1, crlf=False):
    rng = np.random.default_rng(seed)
    Y = rng.normal(10.0, 2.0, size=(n, m))
    eol = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        f.write(",".join(['"ID"'] + [f'"trait{j}"' for j in range(m)] + ['"sex"']) + eol)
        for i in range(n):
            f.write(",".join([f'"BXD{i}"'] + [repr(float(v)) for v in Y[i]] + [str(i % 2)]) + eol)
    return Y


def write_geno(path, n=79, p=41, seed=2):
    rng = np.random.default_rng(seed)
    Pr = np.round(rng.uniform(0, 1, size=(n, p)), 6)
    with open(path, "w") as f:
        f.write(",".join(["id"] + [x for j in range(p) for x in (f"m{j}.B", f"m{j}.D")]) + "\n")
        for i in range(n):
            f.write(",".join([f"BXD{i}"] + [x for j in range(p) for x in (f"{Pr[i, j]:.6f}", f"{1 - Pr[i, j]:.6f}")]) + "\n")
        f.write("\n")  # trailing blank line
    return Pr


@pytest.mark.parametrize("crlf", [False, True])
def test_read_bxd_pheno(tmp_path, crlf):
    path = str(tmp_path / "pheno.csv")
    Y = write_pheno(path, crlf=crlf)
    got = readBXDpheno(path)
    assert got.shape == Y.shape and got.flags["F_CONTIGUOUS"]
    assert np.array_equal(got, Y)  # repr() round-trips exactly, from_chars is correctly rounded
    assert np.array_equal(got, orc.read_bxd_pheno(path))


def test_read_genoprob_excludes_complements(tmp_path):
    path = str(tmp_path / "geno.csv")
    Pr = write_geno(path)
    got = readGenoProb_ExcludeComplements(path)
    assert got.shape == Pr.shape
    assert np.array_equal(got, orc.read_genoprob_exclude_complements(path))
    assert np.allclose(got, Pr, atol=1e-12)
    assert np.array_equal(readBXDgeno(path), orc.read_bxd_geno(path))


def test_general_selection_and_errors(tmp_path):
    path = str(tmp_path / "m.txt")
    with open(path, "w") as f:
        f.write("# comment\n1;2;3;4\n5;6;7;8\n")
    a = read_csv_matrix(path, skip_rows=1, first_col=0, col_step=1, delim=";")
    assert np.array_equal(a, np.array([[1, 2, 3, 4], [5, 6, 7, 8]], dtype=float))
    b = read_csv_matrix(path, skip_rows=1, first_col=1, col_step=2, delim=";")
    assert np.array_equal(b, np.array([[2, 4], [6, 8]], dtype=float))
    with pytest.raises(BlmmError, match="cannot open"):
        read_csv_matrix(str(tmp_path / "missing.csv"))
    bad = str(tmp_path / "bad.csv")
    with open(bad, "w") as f:
        f.write("h1,h2\n1.0,NA\n")
    with pytest.raises(BlmmError, match="non-numeric field at row 2, column 2"):
        read_csv_matrix(bad, skip_rows=1)
    ragged = str(tmp_path / "ragged.csv")
    with open(ragged, "w") as f:
        f.write("h\n1,2,3\n4,5\n")
    with pytest.raises(BlmmError, match="different number of fields"):
        read_csv_matrix(ragged, skip_rows=1)
    longer = str(tmp_path / "longer.csv")
    with open(longer, "w") as f:
        f.write("h\n1,2,3\n4,5,6,7\n")  # MORE fields than the first data row is ragged too
    with pytest.raises(BlmmError, match="different number of fields"):
        read_csv_matrix(longer, skip_rows=1)
    with pytest.raises(BlmmError, match="different number of fields"):
        read_csv_matrix(longer, skip_rows=1, first_col=0, col_step=1, drop_last_cols=1)
