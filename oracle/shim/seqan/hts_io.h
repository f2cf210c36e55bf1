// Oracle build shim (test infrastructure, not product code).
// Shadows the reference's seqan/hts_io.h so that popdel.cpp compiles without htslib.
// Type-check-only stand-ins for the htslib-backed API used by `popdel profile`;
// `popdel call` and `popdel view` never execute any of it.
#ifndef POPDEL_B200_ORACLE_SHIM_HTS_IO_H
#define POPDEL_B200_ORACLE_SHIM_HTS_IO_H
#include <map>
#include <seqan/bam_io.h>
struct kstring_t { size_t l, m; char *s; };
#define KS_INITIALIZE {0,0,NULL}
inline void ks_free(kstring_t *) {}
inline int sam_hdr_find_line_pos(void *, const char *, int, kstring_t *) { return -1; }
enum htsExactFormat { unknown_format, bam, cram };
struct htsFormat { htsExactFormat format; };
struct htsFileStub { htsFormat format; const char * fn_aux; int is_cram; };
namespace seqan {
struct HtsFile { htsFileStub * fp; void * hdr; const char * filename; HtsFile(): fp(NULL), hdr(NULL), filename("") {} };
typedef HtsFile HtsFileIn;
inline bool open(HtsFile &, const char *, const char * = NULL) { return false; }
inline bool loadIndex(HtsFile &, const char *) { return false; }
template <typename T> inline void getContigNames(T &, HtsFile const &) {}
template <typename T> inline void getContigLengths(T &, HtsFile const &) {}
template <typename T> inline void getContigNameToIDMap(T &, HtsFile const &) {}
inline bool setRegion(HtsFile &, const char *, int, int) { return false; }
inline bool seqFreeReadRegion(BamAlignmentRecord &, HtsFile &) { return false; }
}
#endif
