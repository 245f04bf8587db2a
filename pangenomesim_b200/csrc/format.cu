// format.cu -- K3: 2-bit -> ASCII FASTA / MAF text on the device, streamed through pinned buffers.
#include "engine.h"

#include <cstdio>
#include <cstring>

namespace pgs {

int format_genome_name(int64_t genome_index, char *dst, size_t cap)
{
	// reference util.cxx:15-27: "genome" + zero padding to 3 digits + (index + 1); "genomeref"
	// for -1.  The reference aborts for >= 1000 genomes (negative padding length); here the
	// padding is clamped at zero, which leaves every name below 1000 unchanged.
	int len;
	if (genome_index < 0)
		len = std::snprintf(dst, cap, "genomeref");
	else
		len = std::snprintf(dst, cap, "genome%03lld", (long long)(genome_index + 1));
	return len;
}

int write_outputs(Engine &e, const pgs_output_spec *, pgs_output_report *)
{
	return e.fail(PGS_ERR_STATE, "pgs_write_outputs: formatter not built yet");
}

} // namespace pgs
