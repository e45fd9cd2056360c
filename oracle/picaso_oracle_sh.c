/* spherical-harmonics oracle: added with the SH kernels */
