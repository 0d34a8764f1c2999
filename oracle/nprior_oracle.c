/* placeholder translation unit: the neuronized-prior oracle lands here. */
int orc_nprior_placeholder(void) { return 0; }
