// Bulk loader of the embedding BLOBs of a SQLite `chunks` table (the reference
// reads them in a per-row Python loop, src/database_manager.py:35-63).  Host code only: rows are
// stepped through the SQLite C API and copied straight into the caller's (pinned) [N, D] matrix,
// no Python object per row.  The image has libsqlite3.so.0 (Python's own dependency) but not
// sqlite3.h, so the dozen entry points used are declared here and resolved with dlopen.
#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <string>
#include <type_traits>

#include "../../include/anr_b200.h"
#include "anr_internal.h"

namespace {

struct Sqlite {
  // opaque handles
  using db_t = void;
  using stmt_t = void;
  int (*open_v2)(const char*, db_t**, int, const char*) = nullptr;
  int (*close_v2)(db_t*) = nullptr;
  int (*prepare_v2)(db_t*, const char*, int, stmt_t**, const char**) = nullptr;
  int (*step)(stmt_t*) = nullptr;
  int (*finalize)(stmt_t*) = nullptr;
  int (*column_count)(stmt_t*) = nullptr;
  int (*column_type)(stmt_t*, int) = nullptr;
  int (*column_bytes)(stmt_t*, int) = nullptr;
  const void* (*column_blob)(stmt_t*, int) = nullptr;
  long long (*column_int64)(stmt_t*, int) = nullptr;
  const char* (*errmsg)(db_t*) = nullptr;
  bool ok = false;
};

constexpr int kSqliteOk = 0, kSqliteRow = 100, kSqliteDone = 101;
constexpr int kSqliteOpenReadonly = 0x1;
constexpr int kSqliteBlob = 4;

const Sqlite& sqlite() {
  static Sqlite s;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libsqlite3.so.0", "libsqlite3.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return;
    bool all = true;
    auto sym = [&](auto& fn, const char* name) {
      fn = reinterpret_cast<std::decay_t<decltype(fn)>>(dlsym(h, name));
      all = all && fn != nullptr;
    };
    sym(s.open_v2, "sqlite3_open_v2");
    sym(s.close_v2, "sqlite3_close_v2");
    sym(s.prepare_v2, "sqlite3_prepare_v2");
    sym(s.step, "sqlite3_step");
    sym(s.finalize, "sqlite3_finalize");
    sym(s.column_count, "sqlite3_column_count");
    sym(s.column_type, "sqlite3_column_type");
    sym(s.column_bytes, "sqlite3_column_bytes");
    sym(s.column_blob, "sqlite3_column_blob");
    sym(s.column_int64, "sqlite3_column_int64");
    sym(s.errmsg, "sqlite3_errmsg");
    s.ok = all;
  });
  return s;
}

}  // namespace

extern "C" int anr_sqlite_read_blobs(const char* db_path, const char* sql, void* dst,
                                     int64_t row_bytes, int64_t max_rows, int64_t* rowids,
                                     int64_t* n_rows_out, int32_t* uniform_out) {
  if (!db_path || !sql || !dst || !n_rows_out || !uniform_out || row_bytes < 1 || max_rows < 0)
    return anr::set_error(ANR_ERR_INVALID, "anr_sqlite_read_blobs: bad argument");
  *n_rows_out = 0;
  *uniform_out = 0;
  const Sqlite& s = sqlite();
  if (!s.ok) return anr::set_error(ANR_ERR_UNSUPPORTED, "libsqlite3.so.0 could not be loaded");
  Sqlite::db_t* db = nullptr;
  if (s.open_v2(db_path, &db, kSqliteOpenReadonly, nullptr) != kSqliteOk) {
    const std::string msg = db ? s.errmsg(db) : "out of memory";
    if (db) s.close_v2(db);
    return anr::set_error(ANR_ERR_INVALID, "sqlite3_open_v2", msg.c_str());
  }
  Sqlite::stmt_t* stmt = nullptr;
  if (s.prepare_v2(db, sql, -1, &stmt, nullptr) != kSqliteOk || !stmt) {
    const std::string msg = s.errmsg(db);
    if (stmt) s.finalize(stmt);
    s.close_v2(db);
    return anr::set_error(ANR_ERR_INVALID, "sqlite3_prepare_v2", msg.c_str());
  }
  if (s.column_count(stmt) != 2) {
    s.finalize(stmt);
    s.close_v2(db);
    return anr::set_error(ANR_ERR_INVALID,
                          "anr_sqlite_read_blobs: the statement must yield (rowid, blob)");
  }
  unsigned char* out = static_cast<unsigned char*>(dst);
  int64_t n = 0;
  bool uniform = true;
  int rc;
  while ((rc = s.step(stmt)) == kSqliteRow) {
    // a row that is not a BLOB of exactly row_bytes (NULL, text, another width) or a table
    // that grew past max_rows: stop, the caller takes its row-by-row path for the whole table
    if (n >= max_rows || s.column_type(stmt, 1) != kSqliteBlob ||
        s.column_bytes(stmt, 1) != row_bytes) {
      uniform = false;
      break;
    }
    const void* blob = s.column_blob(stmt, 1);
    if (!blob) {
      uniform = false;
      break;
    }
    std::memcpy(out + n * row_bytes, blob, static_cast<size_t>(row_bytes));
    if (rowids) rowids[n] = s.column_int64(stmt, 0);
    ++n;
  }
  int status = ANR_OK;
  if (uniform && rc != kSqliteDone) {
    const std::string msg = s.errmsg(db);
    status = anr::set_error(ANR_ERR_INVALID, "sqlite3_step", msg.c_str());
  }
  s.finalize(stmt);
  s.close_v2(db);
  *n_rows_out = n;
  *uniform_out = uniform ? 1 : 0;
  return status;
}
