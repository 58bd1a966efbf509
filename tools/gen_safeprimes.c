/* tools/gen_safeprimes.c -- regenerates the MWC multiplier table.
 *
 * The per-work-item multiply-with-carry generators of the photon-packet kernel
 * need multipliers a < 2^32 such that p = a*2^32 - 1 is a safe prime
 * (p and (p-1)/2 = a*2^31 - 1 both prime).  The reference ships the first
 * 500 000 such values, searched downwards from a = 4294967118, as
 * xopto/data/primes/safeprimes_base32_500k.npz (found with GMP by
 * xopto/src/rng/find_primes.c:40-93).  This tool recomputes the same sequence
 * from its mathematical definition with a block sieve + deterministic 64-bit
 * Miller-Rabin, so the table in pyxopto_b200/data/ is reproducible without the
 * reference (tools/make_primes_table.py checks both agree).
 *
 * usage: gen_safeprimes N out.bin     (writes N little-endian uint32 values)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((u128)a*b % m); }
static uint64_t powmod(uint64_t b, uint64_t e, uint64_t m) {
	uint64_t r = 1;
	b %= m;
	while (e) { if (e & 1) r = mulmod(r, b, m); b = mulmod(b, b, m); e >>= 1; }
	return r;
}
/* deterministic for all n < 3.3e24 with the first 12 primes as bases */
static int is_prime64(uint64_t n) {
	static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
	if (n < 2) return 0;
	for (int i = 0; i < 12; ++i) {
		if (n == bases[i]) return 1;
		if (n % bases[i] == 0) return 0;
	}
	uint64_t d = n - 1; int s = 0;
	while (!(d & 1)) { d >>= 1; ++s; }
	for (int i = 0; i < 12; ++i) {
		uint64_t x = powmod(bases[i], d, n);
		if (x == 1 || x == n - 1) continue;
		int composite = 1;
		for (int r = 1; r < s; ++r) {
			x = mulmod(x, x, n);
			if (x == n - 1) { composite = 0; break; }
		}
		if (composite) return 0;
	}
	return 1;
}

static uint64_t inv_mod(uint64_t a, uint64_t p) { return powmod(a, p - 2, p); }

#define BLOCK (1u << 22)
#define START 4294967118ull

int main(int argc, char **argv) {
	if (argc < 3) { fprintf(stderr, "usage: %s N out.bin\n", argv[0]); return 2; }
	long n_want = atol(argv[1]);
	FILE *f = fopen(argv[2], "wb");
	if (!f) { perror("open"); return 1; }

	/* odd sieving primes below 2^16 */
	static uint32_t primes[7000]; int np = 0;
	for (uint32_t c = 3; c < 65536; c += 2) {
		int ok = 1;
		for (uint32_t d = 3; d*d <= c; d += 2) if (c % d == 0) { ok = 0; break; }
		if (ok) primes[np++] = c;
	}
	uint8_t *composite = (uint8_t *)malloc(BLOCK);
	long found = 0;
	uint64_t hi = START;                 /* block covers a in (hi - BLOCK, hi] */
	while (found < n_want && hi > BLOCK) {
		uint64_t lo = hi - BLOCK + 1;
		memset(composite, 0, BLOCK);
		for (int i = 0; i < np; ++i) {
			uint64_t p = primes[i];
			/* a*2^32 == 1 (mod p) kills p2; a*2^31 == 1 (mod p) kills p1 */
			uint64_t r2 = inv_mod(powmod(2, 32, p), p);
			uint64_t r1 = inv_mod(powmod(2, 31, p), p);
			uint64_t rs[2] = { r2, r1 };
			for (int k = 0; k < 2; ++k) {
				uint64_t first = lo + ((rs[k] + p - lo % p) % p);
				for (uint64_t a = first; a <= hi; a += p) {
					/* do not strike the prime itself (cannot happen: values are > 2^31) */
					composite[a - lo] = 1;
				}
			}
		}
		for (uint64_t a = hi; a >= lo && found < n_want; --a) {
			if (composite[a - lo]) continue;
			uint64_t p2 = (a << 32) - 1;
			uint64_t p1 = (a << 31) - 1;
			if (is_prime64(p2) && is_prime64(p1)) {
				uint32_t v = (uint32_t)a;
				fwrite(&v, 4, 1, f);
				++found;
			}
		}
		hi = lo - 1;
	}
	fclose(f);
	free(composite);
	fprintf(stderr, "found %ld multipliers, last block ended at a=%llu\n", found,
		(unsigned long long)hi);
	return found == n_want ? 0 : 1;
}
