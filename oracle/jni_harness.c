/* jni_harness.c -- TEST INFRASTRUCTURE ONLY: a minimal fake JNIEnv, so that the reference's UNMODIFIED JNI glue
 * (/root/reference/src/main/native/jni_*.c, compiled where it lies into oracle/_ref/libzstd-jni-b200.so and linked against
 * libzstdb200.so) can be driven without a JVM (none exists in this image).  Only the JNIEnv slots the glue
 * uses are filled (grep "(*env)->" over jni_*.c): byte arrays, direct buffers, long / int fields.  Compiled against the
 * reference's own jni/jni.h.
 *
 *   jh_env()                          -> JNIEnv*
 *   jh_new_array(n) / jh_array_data   -> jbyteArray backed by malloc'ed bytes
 *   jh_new_buffer(addr, cap)          -> a "direct ByteBuffer" (GetDirectBufferAddress / Capacity)
 *   jh_new_object() + jh_get/set_*    -> an object with the fields nativePtr, srcPos, dstPos (long) and consumed, produced (int)
 *   jh_free(obj)
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int kind;                 /* 1 array, 2 direct buffer, 3 object with fields */
    jsize len; jbyte* data;   /* array */
    void* addr; jlong cap;    /* direct buffer */
    jlong longs[4]; jint ints[4];
    int pinned;               /* outstanding GetPrimitiveArrayCritical */
} JhObj;

static int field_index(const char* name) {
    if (!strcmp(name, "nativePtr")) return 0;
    if (!strcmp(name, "srcPos")) return 1;
    if (!strcmp(name, "dstPos")) return 2;
    if (!strcmp(name, "consumed")) return 0;
    if (!strcmp(name, "produced")) return 1;
    return 3;
}
static jclass JNICALL jh_GetObjectClass(JNIEnv* e, jobject o) { (void)e; return (jclass)o; }
static jclass JNICALL jh_FindClass(JNIEnv* e, const char* n) { (void)e; (void)n; return (jclass)0; }
static jfieldID JNICALL jh_GetFieldID(JNIEnv* e, jclass c, const char* name, const char* sig) { (void)e; (void)c; (void)sig; return (jfieldID)(size_t)(field_index(name) + 1); }
static jlong JNICALL jh_GetLongField(JNIEnv* e, jobject o, jfieldID f) { (void)e; return ((JhObj*)o)->longs[(size_t)f - 1]; }
static void JNICALL jh_SetLongField(JNIEnv* e, jobject o, jfieldID f, jlong v) { (void)e; ((JhObj*)o)->longs[(size_t)f - 1] = v; }
static void JNICALL jh_SetIntField(JNIEnv* e, jobject o, jfieldID f, jint v) { (void)e; ((JhObj*)o)->ints[(size_t)f - 1] = v; }
static jsize JNICALL jh_GetArrayLength(JNIEnv* e, jarray a) { (void)e; return ((JhObj*)a)->len; }
static void* JNICALL jh_GetPrimitiveArrayCritical(JNIEnv* e, jarray a, jboolean* isCopy) { (void)e; if (isCopy) *isCopy = JNI_FALSE; ((JhObj*)a)->pinned++; return ((JhObj*)a)->data; }
static void JNICALL jh_ReleasePrimitiveArrayCritical(JNIEnv* e, jarray a, void* p, jint mode) { (void)e; (void)p; (void)mode; ((JhObj*)a)->pinned--; }
static jbyte* JNICALL jh_GetByteArrayElements(JNIEnv* e, jbyteArray a, jboolean* isCopy) { (void)e; if (isCopy) *isCopy = JNI_FALSE; return ((JhObj*)a)->data; }
static void JNICALL jh_ReleaseByteArrayElements(JNIEnv* e, jbyteArray a, jbyte* p, jint mode) { (void)e; (void)a; (void)p; (void)mode; }
static void JNICALL jh_GetByteArrayRegion(JNIEnv* e, jbyteArray a, jsize start, jsize n, jbyte* buf) { (void)e; memcpy(buf, ((JhObj*)a)->data + start, (size_t)n); }
static void* JNICALL jh_GetDirectBufferAddress(JNIEnv* e, jobject b) { (void)e; return b ? ((JhObj*)b)->addr : NULL; }
static jlong JNICALL jh_GetDirectBufferCapacity(JNIEnv* e, jobject b) { (void)e; return b ? ((JhObj*)b)->cap : -1; }
static jint JNICALL jh_ThrowNew(JNIEnv* e, jclass c, const char* msg) { (void)e; (void)c; (void)msg; return 0; }
static void JNICALL jh_DeleteLocalRef(JNIEnv* e, jobject o) { (void)e; (void)o; }
static jstring JNICALL jh_NewStringUTF(JNIEnv* e, const char* s) { (void)e; return (jstring)s; }     /* the C string itself */

static struct JNINativeInterface_ g_table;
static const struct JNINativeInterface_* g_env = NULL;

JNIEnv* jh_env(void) {
    if (!g_env) {
        memset(&g_table, 0, sizeof g_table);
        g_table.GetObjectClass = jh_GetObjectClass; g_table.FindClass = jh_FindClass; g_table.GetFieldID = jh_GetFieldID;
        g_table.GetLongField = jh_GetLongField; g_table.SetLongField = jh_SetLongField; g_table.SetIntField = jh_SetIntField;
        g_table.GetArrayLength = jh_GetArrayLength;
        g_table.GetPrimitiveArrayCritical = jh_GetPrimitiveArrayCritical; g_table.ReleasePrimitiveArrayCritical = jh_ReleasePrimitiveArrayCritical;
        g_table.GetByteArrayElements = jh_GetByteArrayElements; g_table.ReleaseByteArrayElements = jh_ReleaseByteArrayElements;
        g_table.GetByteArrayRegion = jh_GetByteArrayRegion;
        g_table.GetDirectBufferAddress = jh_GetDirectBufferAddress; g_table.GetDirectBufferCapacity = jh_GetDirectBufferCapacity;
        g_table.ThrowNew = jh_ThrowNew; g_table.DeleteLocalRef = jh_DeleteLocalRef; g_table.NewStringUTF = jh_NewStringUTF;
        g_env = &g_table;
    }
    return (JNIEnv*)&g_env;
}
void* jh_new_array(int n) { JhObj* o = (JhObj*)calloc(1, sizeof *o); o->kind = 1; o->len = n; o->data = (jbyte*)calloc((size_t)n + 1, 1); return o; }
void* jh_array_data(void* a) { return ((JhObj*)a)->data; }
int jh_array_pinned(void* a) { return ((JhObj*)a)->pinned; }
void* jh_new_buffer(void* addr, long long cap) { JhObj* o = (JhObj*)calloc(1, sizeof *o); o->kind = 2; o->addr = addr; o->cap = cap; return o; }
void* jh_new_object(void) { JhObj* o = (JhObj*)calloc(1, sizeof *o); o->kind = 3; return o; }
long long jh_get_long(void* o, int i) { return ((JhObj*)o)->longs[i]; }
void jh_set_long(void* o, int i, long long v) { ((JhObj*)o)->longs[i] = v; }
int jh_get_int(void* o, int i) { return ((JhObj*)o)->ints[i]; }
void jh_free(void* p) { JhObj* o = (JhObj*)p; if (!o) return; if (o->kind == 1) free(o->data); free(o); }
