/* placeholder translation unit: the block-sampler oracle lands here. */
int orc_multi_placeholder(void) { return 0; }
