"""CPU: the SQ8 restatement (oracle/sq8_oracle.py) against the reference's own cases
(ScalarQuantizerTests.cs, VectorMathTests.cs:132-155)."""
import numpy as np

from oracle import sq8_oracle as sq


def test_quantize_dequantize_round_trip():  # ScalarQuantizerTests.cs:11-29
    v = np.array([0.0, 0.5, 1.0, -1.0], np.float32)
    q, mn, mx = sq.quantize(v)
    assert len(q) == 4 and mn == np.float32(-1.0) and mx == np.float32(1.0)
    assert np.abs(sq.dequantize(q, mn, mx) - v).max() <= 0.02


def test_quantize_flat_vector():  # :32-45
    q, mn, mx = sq.quantize([0.5, 0.5, 0.5])
    assert mn == mx == np.float32(0.5) and (q == 0).all()
    assert (sq.dequantize(q, mn, mx) == np.float32(0.5)).all()


def test_quantize_span_overload():  # :48-61
    q, mn, mx = sq.quantize([0.0, 1.0])
    assert (mn, mx) == (0.0, 1.0) and q.tolist() == [0, 255]


def test_round_half_to_even():  # Math.Round default (MidpointRounding.ToEven)
    # range 255 -> scale 1: values k + 0.5 round to the even neighbour
    v = np.array([0.0, 0.5, 1.5, 2.5, 254.5, 255.0], np.float32)
    assert sq.quantize(v)[0].tolist() == [0, 0, 2, 2, 254, 255]


def test_8bit_distances_exact():  # VectorMathTests.cs:132-155
    assert sq.l2sq_8bit([10, 20, 255], [12, 18, 250]) == 33
    assert sq.dot_8bit([10, 5, 2], [2, 4, 100]) == 240
