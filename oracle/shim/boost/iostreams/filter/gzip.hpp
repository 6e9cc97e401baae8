// Test-infrastructure shim: tag type standing in for boost::iostreams::gzip_decompressor
// (match.cpp:57). The actual inflate happens in filtering_stream.hpp.
#pragma once
namespace boost { namespace iostreams { struct gzip_decompressor {}; }}
