/* oracle/ref_shim/ref_raw2struct.c -- TEST INFRASTRUCTURE: runs the reference's own xmi_output_raw2struct
 * (src/xmi_data_structs.c:1368-1519, extracted by oracle/build_ref.sh into oracle/_ref/raw2struct.inc at build time; the
 * text is never committed) so that the XMSO writer's history mapping (xmb_output_write_to_xml_file, host_io.cpp) can be
 * pinned against it.  What the function needs from the rest of the reference is supplied here:
 *   - GLib allocation calls: oracle/ref_shim/glib.h (libc);
 *   - xmi_lines[] : the reference's src/xmi_lines.c, compiled alongside;
 *   - xmi_input_copy: an alias (the caller keeps the input alive, nothing is freed through the copy);
 *   - LineEnergy (xraylib): returns 0 -- line energies are xraylib data, not part of the mapping under test;
 *   - xmi_cmp_int, VERSION. */
#include <stdio.h>
#include "xmi_data_structs.h"
#include "xmi_lines.h"
#define VERSION "8.1"
static double LineEnergy(int Z, int line, void *error) { (void)Z; (void)line; (void)error; return 0.0; }
void xmi_input_copy(xmi_input *A, xmi_input **B) { *B = A; }
int xmi_cmp_int(const void *a, const void *b) { return *((const int *)a) - *((const int *)b); }

#include "raw2struct.inc"

/* Flat view for ctypes: one row per (element, line, interaction) of the brute-force (which = 0) or the
 * variance-reduction (which = 1) history.  Returns the number of rows (<= cap written), -1 on error. */
int ref_output_raw2struct_rows(void *input, double *brute_history, double *var_red_history, double **channels_conv,
                               double *channels_unconv, int use_zero_interactions, int which, int cap, int *Z,
                               char *line_type /* [cap][10] */, double *element_total, double *line_total,
                               int *interaction_number, double *counts, double *unconv_checksum) {
	xmi_output *o = xmi_output_raw2struct((xmi_input *)input, brute_history, var_red_history, channels_conv, channels_unconv,
	                                      (char *)"in.xmsi", use_zero_interactions);
	if (!o) return -1;
	const xmi_fluorescence_line_counts *h = which ? o->var_red_history : o->brute_force_history;
	const int nh = which ? o->nvar_red_history : o->nbrute_force_history;
	int n = 0;
	for (int i = 0; i < nh; i++)
		for (int j = 0; j < h[i].n_lines; j++)
			for (int k = 0; k < h[i].lines[j].n_interactions; k++, n++) {
				if (n >= cap) continue;
				Z[n] = h[i].atomic_number;
				strncpy(line_type + (size_t)n * 10, h[i].lines[j].line_type, 9);
				line_type[(size_t)n * 10 + 9] = 0;
				element_total[n] = h[i].total_counts;
				line_total[n] = h[i].lines[j].total_counts;
				interaction_number[n] = h[i].lines[j].interactions[k].interaction_number;
				counts[n] = h[i].lines[j].interactions[k].counts;
			}
	if (unconv_checksum) {   /* sum over the rows the reference copies (row 0 only with use_zero_interactions) */
		double s = 0.0;
		for (int i = 0; i <= o->ninteractions; i++)
			for (int j = 0; j < ((xmi_input *)input)->detector->nchannels; j++) s += o->channels_unconv[i][j] * (double)(i + 1);
		*unconv_checksum = s;
	}
	return n;
}

/* xmi_input_validate (src/xmi_data_structs.c:899-1255), cut out by oracle/build_ref.sh like the function above */
#include <string.h>
#define g_return_val_if_fail(expr, val) do { if (!(expr)) return (val); } while (0)
#include "input_validate.inc"
int ref_input_validate(void *input) { return (int)xmi_input_validate((xmi_input *)input); }
