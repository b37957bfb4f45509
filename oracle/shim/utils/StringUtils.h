// Stand-in (empty) for glare-core utils/StringUtils.h; AudioResampler.cpp includes it but uses nothing from it.
#pragma once
