// Oracle build shim: tabix reading is unused by `popdel call`; this empty guard shadows the
// reference's seqan/vcf_io/tabix.h (which needs htslib, absent in this image).
#ifndef POPDEL_B200_ORACLE_SHIM_TABIX_H
#define POPDEL_B200_ORACLE_SHIM_TABIX_H
#endif
