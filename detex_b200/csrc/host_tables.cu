// Data symbols behind the inline helpers of the reference's public header: a host program compiled against
// reference detex.h that calls detexClamp0To255(), detexDivide0To767By3(), detexDivide0To1791By7() or
// detexDivide0To1279By5() (detex.h:933-982) references these four arrays, so a drop-in library has to export
// them.  The kernels do not use them (min/max and plain division are exact replacements); the
// contents are generated here from their definitions -- clamp.c:23 is clamp(i - 255, 0, 255) for i = 0..766,
// division-tables.c:24,123,350 are i / 3, i / 7, i / 5 -- and compared with the compiled reference's symbols
// by tests/test_library_cpu.py.
#include <cstdint>

namespace {
template <int N> struct Table { uint8_t v[N]; };
template <int N, int D> constexpr Table<N> division_table() {
	Table<N> t{};
	for (int i = 0; i < N; i++)
		t.v[i] = (uint8_t)(i / D);
	return t;
}
constexpr Table<767> clamp_table() {
	Table<767> t{};
	for (int i = 0; i < 767; i++) {
		const int x = i - 255;
		t.v[i] = (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
	}
	return t;
}
}  // namespace

extern "C" {
#define DETEX_DATA __attribute__((visibility("default"))) extern const
// (the arrays are the first and only member of their Table: same address, same layout as uint8_t[N])
DETEX_DATA Table<767> detex_clamp0to255_table;
DETEX_DATA Table<768> detex_division_by_3_table;
DETEX_DATA Table<1792> detex_division_by_7_table;
DETEX_DATA Table<1280> detex_division_by_5_table;
const Table<767> detex_clamp0to255_table = clamp_table();
const Table<768> detex_division_by_3_table = division_table<768, 3>();
const Table<1792> detex_division_by_7_table = division_table<1792, 7>();
const Table<1280> detex_division_by_5_table = division_table<1280, 5>();
}
