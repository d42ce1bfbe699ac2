/* placeholder translation unit; entropic family (KBC / MRTEntropic) restated here. */
int orc_entropic_available(void) { return 0; }
