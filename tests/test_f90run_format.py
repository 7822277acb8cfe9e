"""oracle/f90run's restatement of Fortran format-directed output (rt.fwrite) against records whose text follows from the Fortran 2008
standard (10.7: data edit descriptors; 10.4: reversion) and is what gfortran prints (values that are exact decimal ties, where processors may differ, are avoided).  The comparison of the reference's ocean.stats
record with the oracle's formatter (tests/refcases.py, cases diag/write_energy*) rests on this function."""
import pytest

from oracle.f90run import rt

CASES = [
    # the descriptors of the ocean.stats record (MOM_sum_output.F90:880-905) and of day_str / n_str (:852-859)
    ("(F12.3)", [0.5], "       0.500"),
    ("(F15.3)", [123456789.0627], "  123456789.063"),
    ("(I6)", [48], "    48"),
    ("(ES22.16)", [8.0305275240582147e-02], "8.0305275240582147E-02"),
    ("(F8.5)", [0.00205], " 0.00205"),
    ("(es11.4)", [0.15008], " 1.5008E-01"),
    ("(ES11.5)", [6.73559e17], "6.73559E+17"),
    ("(f8.4)", [35.0626], " 35.0626"),
    ("(f8.4)", [-1.99996], " -2.0000"),
    ("(ES9.2)", [-1.0e-18], "-1.00E-18"),
    ("(ES9.2)", [0.0], " 0.00E+00"),
    ('(A,",",A,",", I6,", En ",ES22.16)', ["     2", "       0.083", 2, 0.25], "     2,       0.083,     2, En 2.5000000000000000E-01"),
    # signs, optional leading zero, overflow
    ("(F5.3)", [-0.5], "-.500"),
    ("(F6.3)", [-0.5], "-0.500"),
    ("(F4.3)", [0.5], ".500"),
    ("(F4.1)", [123.45], "****"),
    ("(I3)", [1000], "***"),
    ("(I5)", [-42], "  -42"),
    ("(I5.3)", [7], "  007"),
    ("(I0)", [12345], "12345"),
    # exponents of three digits drop the letter; E has a zero before the point when it fits
    ("(ES9.2)", [1.0e-100], " 1.00-100"),
    ("(ES12.4E3)", [1.5e10], " 1.5000E+010"),
    ("(E12.4)", [8.03e-2], "  0.8030E-01"),
    ("(E10.4)", [-8.03e-2], "-.8030E-01"),
    ("(E12.4)", [0.0], "  0.0000E+00"),
    # characters, logicals, hexadecimal, positioning
    ("(A5)", ["abc"], "  abc"),
    ("(A2)", ["abcdef"], "ab"),
    ("(A,1X,A)", ["x", "y"], "x y"),
    ("(L2)", [True], " T"),
    ("(Z8)", [255], "      FF"),
    ("(Z16.16)", [1.0], "3FF0000000000000"),
    ("(3X,I2)", [5], "    5"),
    # repeat counts and groups, the colon, reversion
    ('("MOM Date",i7,2("/",i2.2)," ",i2.2,2(":",i2.2))', [1900, 1, 2, 3, 4, 5], "MOM Date   1900/01/02 03:04:05"),
    ("(2I3)", [1, 2], "  1  2"),
    ("(I2)", [1, 2, 3], " 1\n 2\n 3"),
    ('(I2,:,", ")', [1], " 1"),
    ('(I2,", ")', [1], " 1, "),
    ('("x",(I2))', [1, 2], "x 1\n 2"),                       # reversion restarts at the last top-level group
    ("(I2/I2)", [1, 2], " 1\n 2"),
    # not-a-number and infinities
    ("(F8.3)", [float("nan")], "     NaN"),
    ("(ES12.4)", [float("inf")], "    Infinity"),
    ("(F5.1)", [float("-inf")], " -Inf"),
]


@pytest.mark.parametrize("fmt,items,want", CASES)
def test_fwrite(fmt, items, want):
    assert rt.fwrite(fmt, items) == want


def test_unimplemented_descriptors_give_a_marker_not_a_wrong_record():
    assert rt.fwrite("(G12.4)", [1.0]).startswith("<formatted output not reproduced")
