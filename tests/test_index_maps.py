"""CPU: the index identities two kernels rely on (pure integer checks, no library).

* seg_scan_kernel (csrc/air.cu): in the coset-order prefix sum of a bit-reversed circle-domain column
  (simd/prefix_sum.rs:19, index maps core/utils.rs:92-143) position j = sigma * 2^(Q+1) + 2i (+1) lives at row
  bitrev_Q(i) << (s+1) | bitrev_s(sigma) << 1 (even j) and at the bit-complement of that row (odd j).
* twiddle_kernel (csrc/poly.cu): the 4 consecutive entries of a thread differ only in the top 2 bits of the bit-reversed index.
"""
import pytest


def brev(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def coset_pos_to_row(j, L):
    # coset index -> circle-domain index (core/utils.rs:121-143) -> bit-reversed storage row
    cd = ((2 << L) - j) >> 1 if j & 1 else j >> 1
    return brev(cd, L)


@pytest.mark.parametrize("L", [12, 13, 16])
def test_segmented_scan_rows_are_the_coset_order(L):
    Q = 6
    s = L - 1 - Q
    n1 = (1 << L) - 1
    seen = set()
    for tau in range(1 << s):
        sigma = brev(tau, s)
        for i in range(1 << Q):
            row_even = (brev(i, Q) << (s + 1)) | (tau << 1)
            row_odd = n1 - row_even
            j = sigma * (1 << (Q + 1)) + 2 * i
            assert coset_pos_to_row(j, L) == row_even
            assert coset_pos_to_row(j + 1, L) == row_odd
            seen.add(row_even)
            seen.add(row_odd)
    assert len(seen) == 1 << L  # every row belongs to exactly one (segment, step, parity)


@pytest.mark.parametrize("bits", [2, 3, 7, 12])
def test_four_consecutive_twiddle_indices_share_all_but_the_top_two_reversed_bits(bits):
    for base in range(0, 1 << bits, 4):
        shared = brev(base, bits)
        assert shared >> (bits - 2) == 0
        for q in range(4):
            assert brev(base + q, bits) == shared | (brev(q, 2) << (bits - 2))
