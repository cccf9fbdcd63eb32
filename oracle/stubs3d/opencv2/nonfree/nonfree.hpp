/* stub: moped3d util.hpp:56 includes this header; nothing on the path uses it */
