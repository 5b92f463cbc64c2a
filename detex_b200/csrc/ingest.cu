// KTX / DDS ingest: the callers' side of the decode path.
//
//  * Host loaders with the reference's names, signatures, ownership (malloc'd detexTexture +
//    data, released with free()) and error messages: detexLoadKTXFile[WithMipmaps]
//    (reference ktx.c:36-176), detexLoadDDSFile[WithMipmaps] (dds.c:32-143),
//    detexLoadTextureFile[WithMipmaps] (misc.c:98-121).  The file-format tables restate
//    file-info.c:38-170 as data (OpenGL / DXGI identifiers are public enumerations).
//  * detexb200LoadTextureFileToDevice: header parsed on the host, every mip level's payload read
//    straight into ONE pinned buffer, one asynchronous host-to-device copy, then every level is
//    decoded (or, for uncompressed files, converted) on the same stream into one device
//    allocation.  The caller gets device pointers; no pixel ever crosses PCIe.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <strings.h>
#include <vector>
#include <cuda_runtime.h>
#include "launch.h"
#include "../../include/detex.h"
#include "../../include/detex_b200.h"

using namespace detexb200;

namespace {

void fail(const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	set_error_text(buf);
}

// ---- file-format tables (file-info.c:38-170) ------------------------------------------------------
struct FileFormat {
	uint32_t format;	// texture format (compressed) or pixel format (uncompressed)
	int block;		// block width = height: 4 for compressed, 1 for uncompressed
	int gl_internal, gl_format, gl_type;	// KTX identifiers; gl_internal 0 = not in files
	const char *four_cc;	// DDS FourCC; "DX10" = dx10 holds the DXGI format; "" = none
	int dx10;
};

// order matters: look-ups return the first match (file-info.c:171-203, :211-262)
const FileFormat file_formats[] = {
	{ DETEX_PIXEL_FORMAT_RGB8, 1, 0x1907, 0x1907, 0x1401, "", 0 },
	{ DETEX_PIXEL_FORMAT_RGBA8, 1, 0x1908, 0x1908, 0x1401, "DX10", 28 },
	{ DETEX_PIXEL_FORMAT_R8, 1, 0x8229, 0x1903, 0x1401, "DX10", 61 },
	{ DETEX_PIXEL_FORMAT_SIGNED_R8, 1, 0x8F49, 0x1903, 0x1400, "DX10", 63 },
	{ DETEX_PIXEL_FORMAT_RG8, 1, 0x822B, 0x8227, 0x1401, "DX10", 49 },
	{ DETEX_PIXEL_FORMAT_SIGNED_RG8, 1, 0x8F95, 0x8227, 0x1400, "DX10", 51 },
	{ DETEX_PIXEL_FORMAT_R16, 1, 0x822A, 0x1903, 0x1403, "DX10", 56 },
	{ DETEX_PIXEL_FORMAT_SIGNED_R16, 1, 0x8F98, 0x1903, 0x1402, "DX10", 58 },
	{ DETEX_PIXEL_FORMAT_RG16, 1, 0x8226, 0x8227, 0x1403, "DX10", 35 },
	{ DETEX_PIXEL_FORMAT_SIGNED_RG16, 1, 0x8F99, 0x8227, 0x1402, "DX10", 37 },
	{ DETEX_PIXEL_FORMAT_RGB16, 1, 0x8054, 0x1907, 0x1403, "", 0 },
	{ DETEX_PIXEL_FORMAT_RGBA16, 1, 0x805B, 0x8227, 0x1403, "DX10", 11 },
	{ DETEX_PIXEL_FORMAT_FLOAT_R16, 1, 0x822D, 0x1903, 0x140B, "DX10", 54 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RG16, 1, 0x822F, 0x8227, 0x140B, "DX10", 34 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGB16, 1, 0x1907, 0x1907, 0x140B, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGBA16, 1, 0x1908, 0x1908, 0x140B, "DX10", 10 },
	{ DETEX_PIXEL_FORMAT_FLOAT_R32, 1, 0x822E, 0x1903, 0x1406, "DX10", 41 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RG32, 1, 0x8230, 0x8227, 0x1406, "DX10", 16 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGB32, 1, 0x8815, 0x1907, 0x1406, "DX10", 6 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGBA32, 1, 0x8814, 0x1908, 0x1406, "DX10", 2 },
	{ DETEX_PIXEL_FORMAT_A8, 1, 0x1906, 0x1906, 0x1401, "DX10", 65 },
	{ DETEX_TEXTURE_FORMAT_BC1, 4, 0x83F0, 0, 0, "DXT1", 0 },
	{ DETEX_TEXTURE_FORMAT_BC1A, 4, 0x83F1, 0, 0, "", 0 },
	{ DETEX_TEXTURE_FORMAT_BC2, 4, 0x83F2, 0, 0, "DXT3", 0 },
	{ DETEX_TEXTURE_FORMAT_BC3, 4, 0x83F3, 0, 0, "DXT5", 0 },
	{ DETEX_TEXTURE_FORMAT_RGTC1, 4, 0x8DBB, 0, 0, "DX10", 80 },
	{ DETEX_TEXTURE_FORMAT_SIGNED_RGTC1, 4, 0x8DBC, 0, 0, "DX10", 81 },
	{ DETEX_TEXTURE_FORMAT_RGTC2, 4, 0x8DBD, 0, 0, "DX10", 83 },
	{ DETEX_TEXTURE_FORMAT_SIGNED_RGTC2, 4, 0x8DBE, 0, 0, "DX10", 84 },
	{ DETEX_TEXTURE_FORMAT_BPTC_FLOAT, 4, 0x8E8F, 0, 0, "DX10", 95 },
	{ DETEX_TEXTURE_FORMAT_BPTC_SIGNED_FLOAT, 4, 0x8E8E, 0, 0, "DX10", 96 },
	{ DETEX_TEXTURE_FORMAT_BPTC, 4, 0x8E8C, 0, 0, "DX10", 98 },
	{ DETEX_TEXTURE_FORMAT_ETC1, 4, 0x8D64, 0, 0, "", 0 },
	{ DETEX_TEXTURE_FORMAT_ETC2, 4, 0x9274, 0, 0, "", 0 },
	{ DETEX_TEXTURE_FORMAT_ETC2_PUNCHTHROUGH, 4, 0x9275, 0, 0, "", 0 },
	{ DETEX_TEXTURE_FORMAT_ETC2_EAC, 4, 0x9278, 0, 0, "", 0 },
	{ DETEX_TEXTURE_FORMAT_EAC_R11, 4, 0x9270, 0, 0, "", 0 },
	{ DETEX_TEXTURE_FORMAT_EAC_SIGNED_R11, 4, 0x9271, 0, 0, "", 0 },
	{ DETEX_TEXTURE_FORMAT_EAC_RG11, 4, 0x9272, 0, 0, "", 0 },
	{ DETEX_TEXTURE_FORMAT_EAC_SIGNED_RG11, 4, 0x9273, 0, 0, "", 0 },
	// the reference lists ASTC 4x4 as loadable although it has no decoder (detex.h:610, file-info.c:88)
	{ DETEX_TEXTURE_FORMAT_ASTC_4X4, 4, 0x93B0, 0, 0, "DX10", 134 },
	// name-only pseudo formats of the reference's table; they take part in the DDS look-up
	{ DETEX_PIXEL_FORMAT_RGBX8, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_BGRX8, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGBX16, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_BGRX16, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_R16_HDR, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RG16_HDR, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGB16_HDR, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGBA16_HDR, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_R32_HDR, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RG32_HDR, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGB32_HDR, 1, 0, 0, 0, "", 0 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGBA32_HDR, 1, 0, 0, 0, "", 0 },
};
constexpr int N_FILE_FORMATS = (int)(sizeof(file_formats) / sizeof(file_formats[0]));

struct GlSynonym { uint32_t format; int gl_internal, gl_format, gl_type; };
const GlSynonym gl_synonyms[] = {				// file-info.c:120-130
	{ DETEX_PIXEL_FORMAT_RGB8, 0x8051, 0x1907, 0x1401 }, { DETEX_PIXEL_FORMAT_RGBA8, 0x8058, 0x1908, 0x1401 },
	{ DETEX_PIXEL_FORMAT_FLOAT_RGB16, 0x881B, 0x1907, 0x140B }, { DETEX_PIXEL_FORMAT_FLOAT_RGBA16, 0x881A, 0x1908, 0x140B },
	{ DETEX_PIXEL_FORMAT_A8, 0x803C, 0x1906, 0x1401 },
	{ DETEX_TEXTURE_FORMAT_RGTC1, 0x8C70, 0, 0 }, { DETEX_TEXTURE_FORMAT_SIGNED_RGTC1, 0x8C71, 0, 0 },
	{ DETEX_TEXTURE_FORMAT_RGTC2, 0x8C72, 0, 0 }, { DETEX_TEXTURE_FORMAT_SIGNED_RGTC2, 0x8C73, 0, 0 },
};
struct DdsSynonym { uint32_t format; const char *four_cc; int dx10; };
const DdsSynonym dds_synonyms[] = {				// file-info.c:140-167
	{ DETEX_PIXEL_FORMAT_RGBA8, "DX10", 27 }, { DETEX_PIXEL_FORMAT_RGBA8, "DX10", 30 },
	{ DETEX_PIXEL_FORMAT_RG16, "DX10", 36 }, { DETEX_PIXEL_FORMAT_R16, "DX10", 57 },
	{ DETEX_PIXEL_FORMAT_SIGNED_RG16, "DX10", 38 }, { DETEX_PIXEL_FORMAT_SIGNED_R16, "DX10", 59 },
	{ DETEX_PIXEL_FORMAT_RG8, "DX10", 50 }, { DETEX_PIXEL_FORMAT_R8, "DX10", 62 },
	{ DETEX_PIXEL_FORMAT_SIGNED_RG8, "DX10", 52 }, { DETEX_PIXEL_FORMAT_SIGNED_R8, "DX10", 64 },
	{ DETEX_PIXEL_FORMAT_RGBA16, "DX10", 12 },
	{ DETEX_TEXTURE_FORMAT_BC1, "DX10", 70 }, { DETEX_TEXTURE_FORMAT_BC1, "DX10", 71 },
	{ DETEX_TEXTURE_FORMAT_BC2, "DX10", 73 }, { DETEX_TEXTURE_FORMAT_BC2, "DX10", 74 },
	{ DETEX_TEXTURE_FORMAT_BC3, "DX10", 76 }, { DETEX_TEXTURE_FORMAT_BC3, "DX10", 77 },
	{ DETEX_TEXTURE_FORMAT_RGTC1, "DX10", 79 }, { DETEX_TEXTURE_FORMAT_RGTC1, "BC4U", 0 },
	{ DETEX_TEXTURE_FORMAT_SIGNED_RGTC1, "BC4S", 0 },
	{ DETEX_TEXTURE_FORMAT_RGTC2, "DX10", 82 }, { DETEX_TEXTURE_FORMAT_SIGNED_RGTC2, "BC5S", 0 },
	{ DETEX_TEXTURE_FORMAT_BPTC, "DX10", 97 }, { DETEX_TEXTURE_FORMAT_BPTC_FLOAT, "DX10", 94 },
	{ DETEX_TEXTURE_FORMAT_RGTC1, "ATI1", 0 }, { DETEX_TEXTURE_FORMAT_RGTC2, "ATI2", 0 },
};

const FileFormat *by_format(uint32_t format) {			// file-info.c:172-177
	for (int i = 0; i < N_FILE_FORMATS; i++)
		if (file_formats[i].format == format)
			return &file_formats[i];
	return nullptr;
}

const FileFormat *by_gl(int internal, int format, int type) {	// file-info.c:187-203
	for (int i = 0; i < N_FILE_FORMATS; i++) {
		const FileFormat &f = file_formats[i];
		if (f.gl_internal != 0 && f.gl_internal == internal &&
		    (f.gl_format == 0 || (f.gl_format == format && f.gl_type == type)))
			return &f;
	}
	for (const GlSynonym &s : gl_synonyms)
		if (s.gl_internal == internal && (s.gl_format == 0 || (s.gl_format == format && s.gl_type == type)))
			return by_format(s.format);
	return nullptr;
}

// bit masks of the components of an uncompressed format, misc.c:35-71
void component_masks(uint32_t pf, uint64_t m[4]) {
	const int bits = detexGetComponentSize(pf) * 8, n = detexGetNumberOfComponents(pf);
	auto span = [&](int k) { return (bits >= 64 ? ~0ull : ((1ull << bits) - 1ull)) << (k * bits); };
	m[0] = m[1] = m[2] = m[3] = 0;
	if (n == 1 && (pf & DETEX_PIXEL_FORMAT_ALPHA_COMPONENT_BIT)) {
		m[3] = span(0);
		return;
	}
	m[0] = span(0);
	if (n > 1)
		m[1] = span(1);
	if (n > 2)
		m[2] = span(2);
	if (n > 3)
		m[3] = span(3);
	if (pf & DETEX_PIXEL_FORMAT_BGR_COMPONENT_ORDER_BIT) {
		const uint64_t t = m[0];
		m[0] = m[2];
		m[2] = t;
	}
}

const FileFormat *by_dds(const char *four_cc, int dx10, uint32_t pf_flags, int bitcount, const uint32_t mask[4]) {
	const bool is_dx10 = strncmp(four_cc, "DX10", 4) == 0;		// file-info.c:211-262
	for (int i = 0; i < N_FILE_FORMATS; i++) {
		const FileFormat &f = file_formats[i];
		if (is_dx10) {
			if (f.dx10 == dx10)
				return &f;
			continue;
		}
		if (f.four_cc[0] != '\0' && strncmp(f.four_cc, four_cc, 4) == 0)
			return &f;
		if ((pf_flags & 0x40u) && !detexFormatIsCompressed(f.format) && bitcount <= 32) {	// DDPF_RGB: match masks
			uint64_t m[4];
			component_masks(f.format, m);
			if (detexGetPixelSize(f.format) * 8 == bitcount && m[0] == mask[0] && m[1] == mask[1] && m[2] == mask[2] &&
			    ((pf_flags & 0x1u) == 0 || m[3] == mask[3]))
				return &f;
		}
		if ((pf_flags & 0x2u) && bitcount == 8 && f.format == DETEX_PIXEL_FORMAT_A8)	// DDPF_ALPHA
			return &f;
	}
	for (const DdsSynonym &s : dds_synonyms) {
		if (is_dx10) {
			if (s.dx10 == dx10)
				return by_format(s.format);
		}
		else if (s.four_cc[0] != '\0' && strncmp(s.four_cc, four_cc, 4) == 0)
			return by_format(s.format);
	}
	return nullptr;
}

// ---- container parsing: where every mip level's payload sits in the file ---------------------------
struct Level { int width, height, wib, hib; size_t bytes; long offset; };
struct Parsed {
	const FileFormat *info = nullptr;
	int unit_bytes = 0;		// bytes per block (compressed) or per pixel
	std::vector<Level> levels;
};

uint32_t swap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }

int unit_size(uint32_t format) {
	return detexFormatIsCompressed(format) ? (int)detexGetCompressedBlockSize(format) : detexGetPixelSize(format);
}

// Sizes come from an untrusted header: everything is computed in 64 bits and bounded before it is compared with
// the file or turned into an allocation.  A level is accepted when its dimensions are not negative (the tail of a
// mip chain legitimately reaches 0 x 0, ktx.c:129-131) and at most MAX_DIM, and its payload lies inside the file.
constexpr int64_t MAX_DIM = 1 << 20;		// pixels per side (the largest texture any API accepts is 2^15 - 2^16)
bool level_extent(int width, int height, int block, int unit_bytes, int64_t pos, int64_t file_size,
int *wib_out, int *hib_out, uint64_t *bytes_out) {
	if (width < 0 || height < 0 || width > MAX_DIM || height > MAX_DIM || block <= 0 || unit_bytes <= 0 || pos < 0)
		return false;
	const uint64_t wib = ((uint64_t)width + block - 1) / block, hib = ((uint64_t)height + block - 1) / block;
	const uint64_t bytes = wib * hib * (uint64_t)unit_bytes;		// <= 2^40 * 2^6: no overflow
	if (wib * hib > 0xFFFFFFFFull || bytes > (uint64_t)file_size || (uint64_t)pos > (uint64_t)file_size - bytes)
		return false;
	*wib_out = (int)wib;
	*hib_out = (int)hib;
	*bytes_out = bytes;
	return true;
}

// ktx.c:36-176
bool parse_ktx(FILE *f, const char *filename, int max_mipmaps, Parsed &p) {
	const char *who = "detexLoadKTXFileWithMipmaps";
	static const uint8_t magic[12] = { 0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A };
	uint32_t h[16];
	if (fread(h, 1, 64, f) != 64) {
		fail("%s: Error reading file %s", who, filename);
		return false;
	}
	if (memcmp(h, magic, 12) != 0) {
		fail("%s: Couldn't find KTX signature", who);
		return false;
	}
	const bool other_endian = h[3] == 0x01020304u;
	if (other_endian)
		for (int i = 3; i < 16; i++)
			h[i] = swap32(h[i]);
	const int gl_type = (int)h[4], gl_format = (int)h[6], gl_internal = (int)h[7];
	p.info = by_gl(gl_internal, gl_format, gl_type);
	if (p.info == nullptr) {
		fail("%s: Unsupported format in .ktx file (glInternalFormat = 0x%04X)", who, gl_internal);
		return false;
	}
	p.unit_bytes = unit_size(p.info->format);
	const int b = p.info->block;
	int width = (int)h[9], height = (int)h[10];
	const int in_file = (int)h[14];
	const int n_levels = in_file > max_mipmaps ? max_mipmaps : in_file;
	if (in_file < 0 || (int)h[15] < 0) {
		fail("%s: Error reading file %s", who, filename);
		return false;
	}
	int64_t pos = 64 + (int64_t)(int)h[15];			// key/value data is skipped
	fseek(f, 0, SEEK_END);
	const int64_t file_size = ftell(f);
	if (pos > file_size) {
		fail("%s: Error reading file %s", who, filename);
		return false;
	}
	for (int i = 0; i < n_levels; i++) {
		uint32_t image_size;
		if (pos + 4 > file_size || fseek(f, (long)pos, SEEK_SET) != 0 || fread(&image_size, 1, 4, f) != 4) {
			fail("%s: Error reading file %s", who, filename);
			return false;
		}
		if (other_endian)
			image_size = swap32(image_size);
		pos += 4;
		int wib, hib;
		uint64_t bytes;
		if (!level_extent(width, height, b, p.unit_bytes, pos, file_size, &wib, &hib, &bytes)) {
			fail("%s: Error reading file %s", who, filename);
			return false;
		}
		if ((uint64_t)image_size != bytes) {
			fail("%s: Error loading file %s: Image size field of mipmap level %d does not match (%u vs %llu)",
				who, filename, i, (unsigned)image_size, (unsigned long long)bytes);
			return false;
		}
		p.levels.push_back({ width, height, wib, hib, (size_t)bytes, (long)pos });
		pos += (int64_t)bytes;
		width >>= 1;
		height >>= 1;
		if (i + 1 < n_levels) {				// mip padding to a multiple of four
			const int pad = 3 - (int)((image_size + 3u) % 4u);
			if (pos + pad > file_size) {
				fail("%s: Error reading file %s", who, filename);
				return false;
			}
			pos += pad;
		}
	}
	return true;
}

// dds.c:32-143
bool parse_dds(FILE *f, const char *filename, int max_mipmaps, Parsed &p) {
	const char *who = "detexLoadDDSFileWithMipmaps";
	char id[4];
	if (fread(id, 1, 4, f) != 4) {
		fail("%s: Error reading file %s", who, filename);
		return false;
	}
	if (memcmp(id, "DDS ", 4) != 0) {
		fail("%s: Couldn't find DDS signature", who);
		return false;
	}
	uint8_t h[124];
	if (fread(h, 1, 124, f) != 124) {
		fail("%s: Error reading file %s", who, filename);
		return false;
	}
	auto u32 = [&](int off) { uint32_t v; memcpy(&v, h + off, 4); return v; };
	int width = (int)u32(12), height = (int)u32(8);
	const uint32_t pf_flags = u32(76);
	const int bitcount = (int)u32(84);
	const uint32_t mask[4] = { u32(88), u32(92), u32(96), u32(100) };
	char four_cc[5] = { 0, 0, 0, 0, 0 };
	memcpy(four_cc, h + 80, 4);
	uint32_t dx10 = 0;
	long pos = 128;
	if (strncmp(four_cc, "DX10", 4) == 0) {
		uint32_t ext[5];
		if (fread(ext, 1, 20, f) != 20) {
			fail("%s: Error reading file %s", who, filename);
			return false;
		}
		dx10 = ext[0];
		if (ext[1] != 3) {
			fail("%s: Only 2D textures supported for .dds files", who);
			return false;
		}
		pos += 20;
	}
	p.info = by_dds(four_cc, (int)dx10, pf_flags, bitcount, mask);
	if (p.info == nullptr) {
		fail("%s: Unsupported format in .dds file (fourCC = %s, DX10 format = %d).", who, four_cc, (int)dx10);
		return false;
	}
	p.unit_bytes = unit_size(p.info->format);
	const int b = p.info->block;
	int in_file = 1;
	if (u32(4) & 0x20000u)
		in_file = (int)u32(24);
	const int n_levels = in_file > max_mipmaps ? max_mipmaps : in_file;
	fseek(f, 0, SEEK_END);
	const int64_t file_size = ftell(f);
	for (int i = 0; i < n_levels; i++) {
		int wib, hib;
		uint64_t bytes;
		if (!level_extent(width, height, b, p.unit_bytes, pos, file_size, &wib, &hib, &bytes)) {
			fail("%s: Error reading file %s", who, filename);
			return false;
		}
		p.levels.push_back({ width, height, wib, hib, (size_t)bytes, pos });
		pos += (long)bytes;
		width >>= 1;
		height >>= 1;
	}
	return true;
}

enum Container { KTX, DDS, UNKNOWN };
Container container_of(const char *filename) {			// misc.c:100-104
	const size_t n = strlen(filename);
	if (n > 4 && strncasecmp(filename + n - 4, ".ktx", 4) == 0)
		return KTX;
	if (n > 4 && strncasecmp(filename + n - 4, ".dds", 4) == 0)
		return DDS;
	return UNKNOWN;
}

bool open_and_parse(const char *filename, Container c, int max_mipmaps, FILE **f_out, Parsed &p) {
	FILE *f = fopen(filename, "rb");
	if (f == nullptr) {
		fail("%s: Could not open file %s", c == KTX ? "detexLoadKTXFileWithMipmaps" : "detexLoadDDSFileWithMipmaps", filename);
		return false;
	}
	const bool ok = c == KTX ? parse_ktx(f, filename, max_mipmaps, p) : parse_dds(f, filename, max_mipmaps, p);
	if (!ok) {
		fclose(f);
		return false;
	}
	*f_out = f;
	return true;
}

bool read_at(FILE *f, long offset, void *dst, size_t bytes) {
	return fseek(f, offset, SEEK_SET) == 0 && fread(dst, 1, bytes, f) == bytes;
}

// The reference's result: an array of malloc'd textures, each with malloc'd data.
bool load_to_host(const char *filename, Container c, int max_mipmaps, detexTexture ***textures_out, int *nu_levels_out) {
	FILE *f = nullptr;
	Parsed p;
	if (!open_and_parse(filename, c, max_mipmaps, &f, p))
		return false;
	const int n = (int)p.levels.size();
	const char *who = c == KTX ? "detexLoadKTXFileWithMipmaps" : "detexLoadDDSFileWithMipmaps";
	detexTexture **textures = (detexTexture **)calloc((size_t)(n > 0 ? n : 1), sizeof(detexTexture *));
	bool ok = textures != nullptr;
	const char *what = "Out of memory loading";
	for (int i = 0; ok && i < n; i++) {
		const Level &l = p.levels[i];
		detexTexture *t = (detexTexture *)malloc(sizeof(detexTexture));
		if (t == nullptr) {
			ok = false;
			break;
		}
		textures[i] = t;
		t->format = p.info->format;
		t->data = (uint8_t *)malloc(l.bytes ? l.bytes : 1);
		t->width = l.width;
		t->height = l.height;
		t->width_in_blocks = l.wib;
		t->height_in_blocks = l.hib;
		if (t->data == nullptr)
			ok = false;
		else if (!read_at(f, l.offset, t->data, l.bytes)) {
			ok = false;
			what = "Error reading";
		}
	}
	if (!ok) {
		for (int j = 0; textures != nullptr && j < n; j++)
			if (textures[j] != nullptr) {
				free(textures[j]->data);
				free(textures[j]);
			}
		free(textures);
		fclose(f);
		fail("%s: %s file %s", who, what, filename);
		return false;
	}
	fclose(f);
	*nu_levels_out = n;
	*textures_out = textures;
	return true;
}

bool first_level(bool ok, detexTexture **textures, int n, detexTexture **texture_out) {
	if (!ok)
		return false;
	if (n < 1) {		// a file that declares zero mip levels; the reference returns an unset pointer here
		free(textures);
		fail("detexLoadTextureFile: file contains no mipmap level");
		return false;
	}
	*texture_out = textures[0];
	free(textures);
	return true;
}

}  // namespace

extern "C" {

bool detexLoadKTXFileWithMipmaps(const char *filename, int max_mipmaps, detexTexture ***textures_out, int *nu_levels_out) {
	return load_to_host(filename, KTX, max_mipmaps, textures_out, nu_levels_out);
}
bool detexLoadKTXFile(const char *filename, detexTexture **texture_out) {
	detexTexture **t = nullptr;
	int n = 0;
	const bool ok = load_to_host(filename, KTX, 1, &t, &n);		// (t and n are outputs: evaluate the call first)
	return first_level(ok, t, n, texture_out);
}
bool detexLoadDDSFileWithMipmaps(const char *filename, int max_mipmaps, detexTexture ***textures_out, int *nu_levels_out) {
	return load_to_host(filename, DDS, max_mipmaps, textures_out, nu_levels_out);
}
bool detexLoadDDSFile(const char *filename, detexTexture **texture_out) {
	detexTexture **t = nullptr;
	int n = 0;
	const bool ok = load_to_host(filename, DDS, 1, &t, &n);		// (t and n are outputs: evaluate the call first)
	return first_level(ok, t, n, texture_out);
}
bool detexLoadTextureFileWithMipmaps(const char *filename, int max_mipmaps, detexTexture ***textures_out, int *nu_levels_out) {
	const Container c = container_of(filename);
	if (c == UNKNOWN) {
		fail("detexLoadTextureFileWithMipmaps: Do not recognize filename extension");
		return false;
	}
	return load_to_host(filename, c, max_mipmaps, textures_out, nu_levels_out);
}
bool detexLoadTextureFile(const char *filename, detexTexture **texture_out) {
	detexTexture **t = nullptr;
	int n = 0;
	const bool ok = detexLoadTextureFileWithMipmaps(filename, 1, &t, &n);		// (t and n are outputs: evaluate the call first)
	return first_level(ok, t, n, texture_out);
}

// ---- file -> device pixels -------------------------------------------------------------------------
int detexb200LoadTextureFileToDevice(const char *filename, int max_mipmaps, uint32_t pixel_format, int layout,
detexb200DeviceImage *levels_out, int *nu_levels_out, uint32_t *texture_format_out, int *all_valid_out, void *stream) {
	const Container c = container_of(filename);
	if (c == UNKNOWN) {
		fail("detexLoadTextureFileWithMipmaps: Do not recognize filename extension");
		return DETEXB200_EINVAL;
	}
	if (levels_out == nullptr || nu_levels_out == nullptr || max_mipmaps < 1) {
		fail("detexb200LoadTextureFileToDevice: invalid argument");
		return DETEXB200_EINVAL;
	}
	FILE *f = nullptr;
	Parsed p;
	if (!open_and_parse(filename, c, max_mipmaps, &f, p))
		return DETEXB200_EINVAL;
	const uint32_t tf = p.info->format;
	const bool compressed = detexFormatIsCompressed(tf);
	const uint32_t native = tf & 0xFFFFu;
	const uint32_t pf = pixel_format == 0 ? native : pixel_format;
	if (compressed && (tf >> 24) > 19) {
		fclose(f);
		fail("detexb200LoadTextureFileToDevice: no decoder for texture format 0x%08X", tf);
		return DETEXB200_EUNSUPPORTED;
	}
	uint32_t chain[5];
	if (pf != native && detexb200ConversionPlan(native, pf, chain) < 0) {
		fclose(f);
		fail("detexConvertPixels: Unable to find conversion path");
		return DETEXB200_EUNSUPPORTED;
	}
	cudaStream_t s = (cudaStream_t)stream;
	const int n = (int)p.levels.size();
	if (n == 0) {		// a file that declares zero mip levels: nothing to hand back (and nothing to free later)
		fclose(f);
		fail("detexb200LoadTextureFileToDevice: file %s contains no mipmap level", filename);
		return DETEXB200_EINVAL;
	}
	const int ps = detexGetPixelSize(pf);
	// one pinned staging buffer, one device buffer for the payloads, one for the pixels (256-byte aligned levels)
	std::vector<size_t> in_off(n), out_off(n), out_bytes(n);
	size_t in_total = 0, out_total = 0;
	for (int i = 0; i < n; i++) {
		const Level &l = p.levels[i];
		in_off[i] = in_total;
		in_total += (l.bytes + 255) & ~(size_t)255;
		const bool linear = layout == DETEXB200_LAYOUT_LINEAR || !compressed;
		out_bytes[i] = (linear ? (size_t)l.width * l.height : (size_t)l.wib * l.hib * 16) * ps;
		out_off[i] = out_total;
		out_total += (out_bytes[i] + 255) & ~(size_t)255;
	}
	uint8_t *h_in = nullptr, *d_in = nullptr, *d_out = nullptr;
	uint32_t *d_flag = nullptr;
	cudaError_t e;
	auto cleanup = [&]() {
		if (f)
			fclose(f);
		if (h_in)
			cudaFreeHost(h_in);
		if (d_in)
			cudaFree(d_in);
		if (d_out)
			cudaFree(d_out);
		if (d_flag)
			cudaFree(d_flag);
	};
	auto bail = [&](const char *what, cudaError_t err) {
		cleanup();
		fail("%s: CUDA error %d (%s)", what, (int)err, cudaGetErrorString(err));
		return DETEXB200_ECUDA;
	};
	if ((e = cudaMallocHost((void **)&h_in, in_total + 256)) != cudaSuccess)
		return bail("cudaMallocHost", e);
	if ((e = cudaMalloc((void **)&d_in, in_total + 256)) != cudaSuccess ||
	    (e = cudaMalloc((void **)&d_out, out_total + 256)) != cudaSuccess ||
	    (e = cudaMalloc((void **)&d_flag, 4)) != cudaSuccess)
		return bail("cudaMalloc", e);
	for (int i = 0; i < n; i++)
		if (!read_at(f, p.levels[i].offset, h_in + in_off[i], p.levels[i].bytes)) {
			cleanup();
			fail("%s: Error reading file %s", c == KTX ? "detexLoadKTXFileWithMipmaps" : "detexLoadDDSFileWithMipmaps", filename);
			return DETEXB200_EINVAL;
		}
	fclose(f);
	f = nullptr;
	if ((e = cudaMemsetAsync(d_flag, 0, 4, s)) != cudaSuccess ||
	    (e = cudaMemcpyAsync(d_in, h_in, in_total, cudaMemcpyHostToDevice, s)) != cudaSuccess)
		return bail("host to device copy", e);
	int rc = DETEXB200_OK;
	for (int i = 0; i < n && rc == DETEXB200_OK; i++) {
		const Level &l = p.levels[i];
		if ((size_t)l.wib * l.hib == 0)
			continue;
		if (compressed)
			rc = detexb200DecodeDevice(tf, d_in + in_off[i], l.width, l.height, l.wib, l.hib, d_out + out_off[i], pf,
				layout, DETEX_MODE_MASK_ALL, 0, d_flag, stream);
		else
			rc = detexb200ConvertPixelsDevice(d_in + in_off[i], (uint64_t)l.width * l.height, native, d_out + out_off[i], pf, stream);
	}
	uint32_t flag = 0;
	if (rc == DETEXB200_OK) {
		if ((e = cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
		    (e = cudaStreamSynchronize(s)) != cudaSuccess)
			return bail("decode", e);
	}
	else
		cudaStreamSynchronize(s);
	cudaFreeHost(h_in);
	cudaFree(d_in);
	cudaFree(d_flag);
	if (rc != DETEXB200_OK) {
		cudaFree(d_out);
		return rc;
	}
	for (int i = 0; i < n; i++) {
		const Level &l = p.levels[i];
		levels_out[i].d_pixels = d_out + out_off[i];
		levels_out[i].pixel_bytes = out_bytes[i];
		levels_out[i].width = l.width;
		levels_out[i].height = l.height;
		levels_out[i].width_in_blocks = l.wib;
		levels_out[i].height_in_blocks = l.hib;
	}
	*nu_levels_out = n;
	if (texture_format_out)
		*texture_format_out = tf;
	if (all_valid_out)
		*all_valid_out = flag == 0;
	return DETEXB200_OK;
}

// levels[0].d_pixels is the base of the single allocation behind all levels
void detexb200FreeDeviceImages(detexb200DeviceImage *levels, int nu_levels) {
	if (levels != nullptr && nu_levels > 0 && levels[0].d_pixels != nullptr) {
		cudaFree(levels[0].d_pixels);
		for (int i = 0; i < nu_levels; i++)
			levels[i].d_pixels = nullptr;
	}
}

}  // extern "C"
