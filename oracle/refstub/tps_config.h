/* Test infrastructure only: empty stand-in for the autotools-generated tps_config.h
 * (reference: tps_config.h.in) so the reference's per-point physics sources compile
 * in place, unmodified, as the parity oracle.  No feature macro (_GPU_, HAVE_GSL, ...) set. */
#pragma once
