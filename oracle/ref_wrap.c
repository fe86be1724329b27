/* oracle/ref_wrap.c -- test infrastructure only (never linked into the product).
 *
 * Lets the reference's unmodified Main.c drive the reference's own CPU forward:
 * Main.c:54 calls ViT_opencl(); here that name forwards to ViT_seq
 * (ViT_seq.c:402), whose prototype is commented out in ViT_seq.h:5.
 */
#include "Network.h"
void ViT_seq(ImageData *image, Network *networks, float **probabilities);
void ViT_opencl(ImageData *image, Network *networks, float **prb)
{
    ViT_seq(image, networks, prb);
}
