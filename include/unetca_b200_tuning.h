/* unetca_b200_tuning.h — development knobs of libunetca_b200.so.  NOT part of the drop-in ABI (unetca_b200.h).
 *
 * Process-global switches used by the A/B sweeps under tools/ and by the cross-check tests (force the generic
 * tcgen05 tile, disable a kernel family, change a grid heuristic).  They are not re-entrant and not stream-ordered:
 * set them only while no call is in flight.  The product path (insar-unet-ca_b200/*.py) never calls them.
 */
#ifndef UNETCA_B200_TUNING_H
#define UNETCA_B200_TUNING_H
#ifdef __cplusplus
extern "C" {
#endif
void unetca_tc_force_block_n(int n);
void unetca_tc_force_wgrad_narrow(int on);
void unetca_tc_force_no_halo(int on);
void unetca_tc_set_convT_wide(int on);
void unetca_tc_set_first_wgrad_swap(int on); /* 1 (default): first-conv weight gradient with (j, o) on the MMA M side */
void unetca_tc_set_convT_wgrad256(int on); /* 1 (default): ConvTranspose weight gradient in 256 x 256 tiles where Cin % 256 == 0; 0: generic */
void unetca_tc_set_convT_pix(int on);     /* 1 (default): ConvTranspose forward through the dedicated pixels-on-N kernel; 0: generic */
void unetca_tc_force_no_pixn(int on);
void unetca_tc_set_pixn_cluster(int n);
void unetca_tc_force_no_kw(int on);
/* key 0 = pixels per thread-row of an elementwise block, 1 = waves of a reduction grid, 2 = quads per thread-row of
 * se_scale_pool, 3 = shared-memory stream kernels on (default) / off */
void unetca_set_tuning(int key, int value);
#ifdef __cplusplus
}
#endif
#endif
