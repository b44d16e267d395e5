// cb_resolver.cuh -- backend identity for the resolver side of the path:
// srvKey(), lib/resolver.js:1157-1171 = base64(sha1(name || "||" || port || "||" || normalised address)),
// for a whole SRV/A result set at once (4096 backends on config C4/C5): one thread per backend.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

constexpr int SRVKEY_LEN = 28; // base64 of a 20-byte SHA-1 digest

__device__ __forceinline__ unsigned int sha1_rol(unsigned int x, int n) { return (x << n) | (x >> (32 - n)); }

// byte i of the padded message (FIPS 180-4 5.1.1): the message, 0x80, zeros, the 64-bit big-endian bit length
__device__ __forceinline__ unsigned int sha1_padded_byte(const unsigned char *m, unsigned int len, unsigned int total,
    unsigned int i) {
	if (i < len)
		return m[i];
	if (i == len)
		return 0x80u;
	if (i >= total - 8) {
		const unsigned long long bits = (unsigned long long)len * 8ull;
		return (unsigned int)((bits >> (8 * (total - 1 - i))) & 0xffull);
	}
	return 0u;
}

__global__ void srv_key_kernel(const unsigned char *msgs, const unsigned int *offs, int n, char *keys) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n)
		return;
	const unsigned char *m = msgs + offs[k];
	const unsigned int len = offs[k + 1] - offs[k];
	const unsigned int total = ((len + 8) / 64 + 1) * 64; // room for 0x80 and the length
	unsigned int h0 = 0x67452301u, h1 = 0xEFCDAB89u, h2 = 0x98BADCFEu, h3 = 0x10325476u, h4 = 0xC3D2E1F0u;
	for (unsigned int blk = 0; blk < total; blk += 64) {
		unsigned int w[16];
#pragma unroll
		for (int t = 0; t < 16; t++) {
			const unsigned int i = blk + 4 * t;
			w[t] = (sha1_padded_byte(m, len, total, i) << 24) | (sha1_padded_byte(m, len, total, i + 1) << 16) |
			    (sha1_padded_byte(m, len, total, i + 2) << 8) | sha1_padded_byte(m, len, total, i + 3);
		}
		unsigned int a = h0, b = h1, c = h2, d = h3, e = h4;
#pragma unroll 1
		for (int t = 0; t < 80; t++) {
			unsigned int wt;
			if (t < 16) {
				wt = w[t];
			} else {
				wt = sha1_rol(w[(t - 3) & 15] ^ w[(t - 8) & 15] ^ w[(t - 14) & 15] ^ w[t & 15], 1);
				w[t & 15] = wt;
			}
			unsigned int f, kk;
			if (t < 20) {
				f = (b & c) | (~b & d);
				kk = 0x5A827999u;
			} else if (t < 40) {
				f = b ^ c ^ d;
				kk = 0x6ED9EBA1u;
			} else if (t < 60) {
				f = (b & c) | (b & d) | (c & d);
				kk = 0x8F1BBCDCu;
			} else {
				f = b ^ c ^ d;
				kk = 0xCA62C1D6u;
			}
			const unsigned int tmp = sha1_rol(a, 5) + f + e + kk + wt;
			e = d;
			d = c;
			c = sha1_rol(b, 30);
			b = a;
			a = tmp;
		}
		h0 += a;
		h1 += b;
		h2 += c;
		h3 += d;
		h4 += e;
	}
	unsigned char dg[21];
	const unsigned int hs[5] = {h0, h1, h2, h3, h4};
#pragma unroll
	for (int j = 0; j < 5; j++) {
		dg[4 * j] = (unsigned char)(hs[j] >> 24);
		dg[4 * j + 1] = (unsigned char)(hs[j] >> 16);
		dg[4 * j + 2] = (unsigned char)(hs[j] >> 8);
		dg[4 * j + 3] = (unsigned char)hs[j];
	}
	dg[20] = 0;
	const char *A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
	char *o = keys + (size_t)k * SRVKEY_LEN;
#pragma unroll
	for (int j = 0; j < 7; j++) { // 7 groups of 3 bytes (the last holds 2 bytes + padding)
		const unsigned int v = ((unsigned int)dg[3 * j] << 16) | ((unsigned int)dg[3 * j + 1] << 8) | (unsigned int)dg[3 * j + 2];
		o[4 * j] = A[(v >> 18) & 63];
		o[4 * j + 1] = A[(v >> 12) & 63];
		o[4 * j + 2] = A[(v >> 6) & 63];
		o[4 * j + 3] = A[v & 63];
	}
	o[27] = '='; // 20 = 6 * 3 + 2: one pad character
}

} // namespace cb
