"""fgq_math.h states that a / b for the LITERAL divisors of the kernels (m 2^k with a small odd m) may be formed as
q = a r; rem = fma(-b, q, a); fma(r, rem, q) with r = RN(1 / b) — the last three operations of the IEEE quotient's fast
path — and still be the correctly rounded quotient for EVERY a in range (FGQ_DIV_LIT, fgq_div_by).  Checked here on the
CPU, in C with the same operations: random dividends over many binades and adversarial ones (the neighbours of b times
odd multiples of half an ulp, where a quotient comes closest to a rounding boundary)."""
import os
import subprocess
import textwrap

SRC = textwrap.dedent(r"""
    #include <math.h>
    #include <stdint.h>
    #include <stdio.h>
    #include <string.h>
    static uint64_t s = 0x9E3779B97F4A7C15ull;
    static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
    static double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
    static int check(double a, double b, double r) {
        const double q = a * r;
        const double rem = fma(-b, q, a);
        const double res = fma(r, rem, q);
        return res == a / b;
    }
    int main(void) {
        const double divs[] = {3, 5, 6, 12, 24, 45, 720, 1152, 2835, 42525, 90720, 181440, 414720, 1403325, 6967296, 10886400};
        long bad = 0, total = 0;
        for (unsigned d = 0; d < sizeof divs / sizeof divs[0]; ++d) {
            const double b = divs[d], r = 1.0 / b;
            for (long i = 0; i < 4000000; ++i) {   /* random significands, exponents within +-200 */
                const uint64_t u = rnd();
                const uint64_t e = 1023 - 200 + (rnd() % 401);
                const double a = from_bits((u & 0x800FFFFFFFFFFFFFull) | (e << 52));
                bad += !check(a, b, r); ++total;
            }
            for (long i = 0; i < 2000000; ++i) {   /* a next to b * (q + half an ulp of q): quotients next to midpoints */
                const uint64_t u = rnd();
                const double q = from_bits((u & 0x000FFFFFFFFFFFFFull) | (1023ull << 52));
                double a = fma(b, q, ldexp(b, -53));              /* RN(b (q + ulp(q)/2)): a / b lands next to a midpoint */
                for (int k = -2; k <= 2; ++k) {
                    double ak = a;
                    for (int j = 0; j < (k < 0 ? -k : k); ++j) ak = nextafter(ak, k < 0 ? 0.0 : INFINITY);
                    bad += !check(ak, b, r); ++total;
                }
            }
        }
        printf("%ld %ld\n", bad, total);
        return 0;
    }
""")


def test_literal_divisors_round_like_ieee_division(tmp_path):
    src = tmp_path / "divlit.c"
    exe = tmp_path / "divlit"
    src.write_text(SRC)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(src), "-lm"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True, timeout=600).stdout.split()
    bad, total = int(out[0]), int(out[1])
    assert total > 2e8 and bad == 0, f"{bad} of {total} quotients differ from a / b"


def test_every_literal_divisor_in_the_kernels_is_covered():
    """The divisors passed to FGQ_DIV_LIT / the R... constants in the sources are all in the list checked above."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "fastgaussquadrature.jl_b200", "csrc")
    found = set()
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh", ".h")):
            txt = open(os.path.join(csrc, f)).read()
            found |= {float(m) for m in re.findall(r"FGQ_DIV_LIT\([^;]*?,\s*([0-9.]+)\)", txt)}
            found |= {float(m) for m in re.findall(r"fgq_div_by_inrange\([^;]*?,\s*([0-9.]+),\s*R[0-9]+\)", txt)}
    assert found, "no literal divisors found: the pattern drifted"
    checked = {3, 5, 6, 12, 24, 45, 720, 1152, 2835, 42525, 90720, 181440, 414720, 1403325, 6967296, 10886400}
    assert found <= checked, found - checked
