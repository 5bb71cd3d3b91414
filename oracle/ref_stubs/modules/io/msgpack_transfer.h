// Shadows modules/io/msgpack_transfer.h (test infrastructure): no msgpack here; the leaf sources compiled into
// oracle/_ref never (de)serialise through it.
#pragma once
#include "modules/io/transfer_object.h"
#include <stdexcept>
template <class T> std::string msgpack_serialize(const T&) { throw std::logic_error("oracle/_ref: msgpack is not available"); }
template <class T> void msgpack_deserialize(T&, const std::string&) { throw std::logic_error("oracle/_ref: msgpack is not available"); }
