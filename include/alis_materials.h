/*
 * alis_materials.h -- material table and colour maps of alis-b200.
 *
 * DATA, kept numerically identical to the reference's material database (reference
 * src/alis_m.h:18-66; fits documented in reference doc/DFFit.ipynb) because drivers name these
 * globals (`putObjects(W, Au, ...)`) and parity with the reference requires the same pole sets.
 * Model:  eps(w) = eps_inf - sum_D wD^2/(w^2 + i gD w) + sum_CP (f w0^2 - i g w0 w)/(w0^2 - (w + i g0)^2),
 * energies in eV (res.spec.eV = 1240 maps them to wavelengths in nm).
 *   { w, gamma }        Drude pole            { 0, sigma }          conductivity
 *   { w, gamma, f }     Lorentz pole          { w, gamma, f, g }    critical-point pole
 * Sources of the optical constants: CRC Handbook of Chemistry & Physics (2006); Johnson & Christy,
 * PRB 6, 4370 (1972) (Ag); Ninomiya & Adachi, JAP 78, 1183 (1995) (CdS).
 *
 * Like the reference header this file DEFINES globals: include it (via alis.h) in exactly one
 * translation unit.  -DNUV selects the 250-750 nm fits instead of the 400-2000 nm ones.
 */
#ifndef ALIS_MATERIALS_H
#define ALIS_MATERIALS_H

matter PM     = { .e = { 0 } };
matter Air    = { .e = { 1 } };
matter Vacuum = { .e = { 1 } };

#ifdef NUV   /* 250-750 nm */
matter Ag = {
	.e = { 1.9201 },
	.P = {
		{ 9.3937, 0.0302 },	/* Drude */
		{ 4.2333, 1.2373, 1.3399, 1.6141 },	/* critical point */
		{ 3.9548, 0.259, -0.1356, 0.2961 },	/* critical point */
		{ 0.8394, 0.5873, 2.652, 6.1529 },	/* critical point */
	}
};
matter Au = {
	.e = { 2.03 },
	.P = {
		{ 8.4306, 0.0 },	/* Drude */
		{ 4.3848, 1.2748, -0.2723, -3.7849 },	/* critical point */
		{ 4.7338, 2.0464, 5.3396, 4.7241 },	/* critical point */
		{ 2.6147, 0.3187, 0.371, 0.723 },	/* critical point */
	}
};
matter Cu = {
	.e = { 1.7411 },
	.P = {
		{ 8.8195, 0.0835 },	/* Drude */
		{ 4.6887, 0.7044, 0.3576, 0.4646 },	/* critical point */
		{ 3.4555, 2.8361, 7.7571, -0.1944 },	/* critical point */
		{ 2.1607, 0.2625, 0.0964, 1.0257 },	/* critical point */
	}
};
matter Al = {
	.e = { 1.0 },
	.P = {
		{ 13.4862, 0.1406 },	/* Drude */
		{ 2.24, 1.4861, 3.8125, -0.5311 },	/* critical point */
		{ 1.8272, 0.3877, 3.8946, 2.3493 },	/* critical point */
		{ 1.5717, 0.2272, 9.7451, -1.7878 },	/* critical point */
	}
};
matter Cr = {
	.e = { 1.0009 },
	.P = {
		{ 9.1925, 0.7681 },	/* Drude */
		{ 3.4893, 0.4922, 0.7387, -0.0827 },	/* critical point */
		{ 0.9432, 0.2803, -21.3216, 29.9879 },	/* critical point */
		{ 2.4741, 0.6165, 2.6444, -5.3032 },	/* critical point */
	}
};
matter Ni = {
	.e = { 2.04 },
	.P = {
		{ 7.8334, 0.5169 },	/* Drude */
		{ 4.3024, 1.0858, 2.3911, 1.602 },	/* critical point */
		{ 1.3054, 0.4506, 8.9446, 6.5275 },	/* critical point */
		{ 3.3426, 1.7382, 1.3531, -3.329 },	/* critical point */
	}
};
matter Ti = {
	.e = { 1.0 },
	.P = {
		{ 12.3768, 5.5499 },	/* Drude */
		{ 3.9759, 0.8038, -1.0395, -0.3514 },	/* critical point */
		{ 2.1995, 1.048, -3.7293, 0.9582 },	/* critical point */
		{ 1.0376, 0.2488, -0.596, 6.2558 },	/* critical point */
	}
};
matter Pd = {
	.e = { 1.7139 },
	.P = {
		{ 9.3495, 0.7156 },	/* Drude */
		{ 4.7399, 0.7234, 0.2268, 0.0691 },	/* critical point */
		{ 2.0356, 3.4652, 6.1216, -2.2174 },	/* critical point */
		{ 1.3945, 0.7049, 7.316, 5.9221 },	/* critical point */
	}
};
matter Pt = {
	.e = { 1.9043 },
	.P = {
		{ 11.5026, 1.6472 },	/* Drude */
		{ 1.2596, 1.3057, 5.818, 14.2589 },	/* critical point */
	}
};
matter Ge = {
	.e = { 1.0 },
	.P = {
		{ 3.3888, 1.0394, 6.1945, 3.7488 },	/* critical point */
		{ 4.0446, 1.7946, 5.5177, -2.0495 },	/* critical point */
		{ 4.3299, 0.2043, 1.1731, -0.9613 },	/* critical point */
		{ 2.269, 0.2808, 3.7623, 1.9046 },	/* critical point */
		{ 2.1174, 0.0858, 0.5609, 0.0901 },	/* critical point */
	}
};
matter Si = {
	.e = { 1.0 },
	.P = {
		{ 4.2785, 0.2217, 2.9926, -0.7684 },	/* critical point */
		{ 5.6927, 1.5593, 2.3925, -1.2208 },	/* critical point */
		{ 3.6614, 0.4785, 4.1976, 2.2951 },	/* critical point */
		{ 3.3756, 0.1172, 1.3635, 1.0095 },	/* critical point */
	}
};
matter aSi = {
	.e = { 1.0 },
	.P = {
		{ 2.9963, 0.7273, 2.1465, 1.9077 },	/* critical point */
		{ 2.083, 1.0256, 0.434, 5.1983 },	/* critical point */
		{ 4.3391, 1.9704, 3.9252, -1.6418 },	/* critical point */
		{ 3.5877, 1.0243, 5.027, -1.4322 },	/* critical point */
	}
};
#else        /* 400-2000 nm */
matter Ag = {
	.e = { 4.0323 },
	.P = {
		{ 9.1864, 0.021 },	/* Drude */
	}
};
matter Au = {
	.e = { 2.1452 },
	.P = {
		{ 8.4146, 0.0255 },	/* Drude */
		{ 2.6203, 0.2147, 0.2082, 0.3059 },	/* critical point */
		{ 3.1543, 1.1662, 4.1353, 2.5005 },	/* critical point */
	}
};
matter Cu = {
	.e = { 4.0934 },
	.P = {
		{ 9.0117, 0.0943 },	/* Drude */
		{ 3.1915, 1.534, 3.6969, -0.0729 },	/* critical point */
		{ 2.1831, 0.2548, 0.2499, 0.914 },	/* critical point */
	}
};
matter Al = {
	.e = { 1.0 },
	.P = {
		{ 13.3237, 0.1128 },	/* Drude */
		{ 1.492, 0.2014, 8.2647, 4.9317 },	/* critical point */
		{ 0.0844, 1.2677, 23.8069, -4.9835 },	/* critical point */
		{ 1.2054, 1.0353, 21.7077, -5.7556 },	/* critical point */
	}
};
matter Cr = {
	.e = { 1.0 },
	.P = {
		{ 4.9087, 0.0465 },	/* Drude */
		{ 2.4765, 0.4536, 1.5602, -2.1909 },	/* critical point */
		{ 0.3726, 0.1348, 21.4832, 36.8394 },	/* critical point */
		{ 1.2778, 1.0631, 49.801, 4.9223 },	/* critical point */
	}
};
matter Ni = {
	.e = { 1.0 },
	.P = {
		{ 5.5189, 0.0658 },	/* Drude */
		{ 1.6545, 0.6339, 0.6792, -3.2872 },	/* critical point */
		{ 0.3545, 0.1956, 57.2374, 79.7729 },	/* critical point */
	}
};
matter Ti = {
	.e = { 1.28 },
	.P = {
		{ 12.6869, 4.8641 },	/* Drude */
		{ 0.4422, 0.0729, -19.4166, -3.9574 },	/* critical point */
		{ 1.1044, 0.9335, -12.1464, 8.0836 },	/* critical point */
		{ 1.5619, 0.5214, 3.0698, -3.7725 },	/* critical point */
	}
};
matter Pd = {
	.e = { 1.5512 },
	.P = {
		{ 6.049, 0.0379 },	/* Drude */
		{ 0.3874, 0.2534, 87.1642, 5.9413 },	/* critical point */
		{ 0.6749, 1.6453, 87.1631, 5.9296 },	/* critical point */
		{ 0.6006, 0.7056, 87.1633, 5.9317 },	/* critical point */
	}
};
matter Pt = {
	.e = { 1.0513 },
	.P = {
		{ 6.5658, 0.2092 },	/* Drude */
		{ 1.1371, 1.1336, 13.527, -0.9232 },	/* critical point */
		{ 0.6331, 0.1574, 0.3107, 19.5815 },	/* critical point */
		{ 0.7853, 0.4354, 15.1059, -0.6342 },	/* critical point */
		{ 0.4794, 0.0923, -5.5467, 26.6143 },	/* critical point */
	}
};
matter Ge = {
	.e = { 1.6259 },
	.P = {
		{ 2.0327, 0.2169, -0.6161, 2.889 },	/* critical point */
		{ 3.5657, 0.5248, 5.8672, 1.3512 },	/* critical point */
		{ 2.097, 0.6267, 10.4417, 2.3849 },	/* critical point */
	}
};
matter Si = {
	.e = { 3.7334 },
	.P = {
		{ 3.6745, 0.1452, 8.0789, 0.6725 },	/* critical point */
	}
};
matter aSi = {
	.e = { 1.6976 },
	.P = {
		{ 3.59, 1.087, 11.3535, -2.5534 },	/* critical point */
		{ 2.1662, 1.1161, -1.0961, 5.9546 },	/* critical point */
	}
};
#endif
/* 400-1040 nm (ordinary / extraordinary CdS) */
matter oCdS = {
	.e = { 1.0 },
	.P = {
		{ 5.0323, 0.7745, 2.3657, 0.1689 },	/* critical point */
		{ 2.7075, 0.4865, 0.2835, 0.4341 },	/* critical point */
		{ 2.4458, 0.0664, 0.0553, 0.0341 },	/* critical point */
		{ 6.1838, 7.2264, 1.4749, -0.1979 },	/* critical point */
	}
};
matter eCdS = {
	.e = { 1.0 },
	.P = {
		{ 4.6077, 0.7854, 2.5833, 0.4566 },	/* critical point */
		{ 2.4621, 0.0784, 0.0674, 0.0377 },	/* critical point */
		{ 2.7499, 0.4034, 0.2438, 0.2882 },	/* critical point */
		{ 5.2255, 7.1767, 1.43, 0.1797 },	/* critical point */
	}
};

/* colour maps: { number of stops (negative: symmetric about zero), 0xRRGGBB ... } */
colormap gray = { 2,
	0xFFFFFF, 0x000000
};
colormap red = { 3,
	0xFFFFFF, 0xFF0000, 0x000000
};
colormap blue = { 3,
	0xFFFFFF, 0x0000FF, 0x000000
};
colormap dkbr = { -5,
	0x002244, 0x0022FF, 0xFFFFFF, 0xFF2200, 0x442200
};
colormap dkrb = { -5,
	0x442200, 0xFF2200, 0xFFFFFF, 0x0022FF, 0x002244
};
colormap hsv = { 7,
	0xFF0000, 0xFFFF00, 0x00FF00, 0x00FFFF, 0x0000FF, 0xFF00FF, 0xFF0000
};
colormap hot = { 9,
	0x000000, 0x550000, 0xAA0000, 0xFF0000, 0xFF5500, 0xFFAA00, 0xFFFF00, 0xFFFF80, 0xFFFFFF
};
colormap jet = { 16,
	0x0000C0, 0x0000FF, 0x0040FF, 0x0080FF, 0x00C0FF, 0x00FFFF, 0x40FFFF, 0x80FFC0, 0xC0FF80, 0xFFFF40, 0xFFFF00, 0xFFC000, 0xFF8000, 0xFF4000, 0xFF0000, 0xC00000
};

#endif /* ALIS_MATERIALS_H */
