// TEST INFRASTRUCTURE ONLY — stand-in for the author's out-of-tree `freon` matrix library
// (swegl Makefile:12,19 expects it at ../freon; it is not vendored and has no pinned version).
// swegl only needs: element access m[r][c], an initializer-list constructor, a 4x4 identity
// and operator* for the node-hierarchy product (vertex_shaders.hpp:18).  The product order
// below (k ascending, accumulator starting at 0) is this repo's DEFINITION of that product;
// the product path keeps it on the host (swegl_b200/host, swegl_b200/scene.py) so the device
// never depends on it.  swegl also relies on freon's transitive standard includes.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <string>
#include <tuple>
#include <vector>

namespace freon
{
template <typename T, size_t R, size_t C>
struct Matrix
{
	T d[R][C];
	Matrix() { for (size_t i = 0; i < R; i++) for (size_t j = 0; j < C; j++) d[i][j] = T(0); }
	Matrix(std::initializer_list<T> l)
	{
		size_t n = 0;
		for (size_t i = 0; i < R; i++) for (size_t j = 0; j < C; j++) d[i][j] = T(0);
		for (const T & v : l) { if (n >= R * C) break; d[n / C][n % C] = v; n++; }
	}
	T * operator[](size_t r) { return d[r]; }
	const T * operator[](size_t r) const { return d[r]; }
};

template <typename T, size_t R, size_t K, size_t C>
Matrix<T, R, C> operator*(const Matrix<T, R, K> & a, const Matrix<T, K, C> & b)
{
	Matrix<T, R, C> out;
	for (size_t i = 0; i < R; i++)
		for (size_t j = 0; j < C; j++)
		{
			T s = T(0);
			for (size_t k = 0; k < K; k++)
				s += a[i][k] * b[k][j];
			out[i][j] = s;
		}
	return out;
}

template <typename T>
struct MatrixIdentity
{
	static inline const Matrix<T, 4, 4> _4{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
};
} // namespace freon
