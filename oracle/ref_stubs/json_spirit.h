// Stand-in: json_spirit is absent; the sources compiled into oracle/_ref only name its namespace (test infrastructure).
#pragma once
namespace json_spirit {}
